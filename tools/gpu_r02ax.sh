#!/bin/bash
# the default bench as the driver launches it at N = 8
set -x
mkdir -p gpurun_out
free -g > gpurun_out/r02ax_free.txt; nproc >> gpurun_out/r02ax_free.txt; nvidia-smi -L >> gpurun_out/r02ax_free.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 3 ) > gpurun_out/r02ax_bench1_8gpu.json 2> gpurun_out/r02ax_bench1_8gpu.err; echo rc=$?
