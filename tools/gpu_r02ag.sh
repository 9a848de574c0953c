#!/bin/bash
# final build on 2 GPUs: the NCCL parity test, and the bench lines of configs 1, 2, 3 under torchrun
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02ag_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02ag_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ag_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29511 bench.py --gpus 2 --config 3 --steps 2 --warmup 3 > gpurun_out/r02ag_bench3_2gpu.json 2> gpurun_out/r02ag_bench3_2gpu.err; echo rc=$?
$T --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02ag_bench1_2gpu.json 2> gpurun_out/r02ag_bench1_2gpu.err; echo rc=$?
$T --master-port 29513 bench.py --gpus 2 --config 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02ag_bench2_2gpu.json 2> gpurun_out/r02ag_bench2_2gpu.err; echo rc=$?
$T --master-port 29514 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/r02ag_ref_2gpu.json 2> gpurun_out/r02ag_ref_2gpu.err; echo rc=$?
