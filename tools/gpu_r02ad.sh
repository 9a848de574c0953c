#!/bin/bash
set -x
mkdir -p gpurun_out
export CB200_COOP=0 CB200_FROBPERS=0
echo "all-shared batch=16576" >> gpurun_out/r02ad_bench.log
CB200_HYBRID=0 timeout 900 python bench.py --config 2 --frob-calls 2 --batch 16576 --terms 64 --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02ad_bench.log 2>&1
echo "hybrid192 batch=28416" >> gpurun_out/r02ad_bench.log
CB200_HYBRID=1 timeout 900 python bench.py --config 2 --frob-calls 2 --batch 28416 --terms 64 --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02ad_bench.log 2>&1
echo "pers208 nosync batch=30784" >> gpurun_out/r02ad_bench.log
CB200_FROBPERS=1 timeout 900 python bench.py --config 2 --frob-calls 2 --batch 30784 --terms 64 --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02ad_bench.log 2>&1
