#!/bin/bash
# two GPUs: the NCCL test of the sharded solve + device-resident gather, and bench.py under torchrun (configs 3 and 1)
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -15 gpurun_out/r02f_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 3 --steps 2 --warmup 3 > gpurun_out/r02f_bench3_2gpu.json 2> gpurun_out/r02f_bench3_2gpu.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02f_bench1_2gpu.json 2> gpurun_out/r02f_bench1_2gpu.err; echo rc=$?
tail -c 600 gpurun_out/r02f_bench3_2gpu.json; tail -n 5 gpurun_out/r02f_bench3_2gpu.err; tail -n 5 gpurun_out/r02f_bench1_2gpu.err
