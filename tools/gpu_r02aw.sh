#!/bin/bash
set -x
mkdir -p gpurun_out
echo "config 4 full size, two lanes" >> gpurun_out/r02aw_bench.log
CB200_COOP=2 timeout 900 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02aw_bench.log 2>&1
echo "config 4 full size, default" >> gpurun_out/r02aw_bench.log
timeout 900 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02aw_bench.log 2>&1
