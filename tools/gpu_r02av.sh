#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "two_lanes or horner" > gpurun_out/r02av_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02av_pytest.log
timeout 900 python tools/quick_coop.py > gpurun_out/r02av_quick.log 2>&1
echo "config 4 two lanes, 200000 points, 500 terms" >> gpurun_out/r02av_bench.log
CB200_COOP=2 timeout 900 python bench.py --config 4 --batch 200000 --terms 500 --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02av_bench.log 2>&1
echo "config 4 default, 200000 points, 500 terms" >> gpurun_out/r02av_bench.log
timeout 900 python bench.py --config 4 --batch 200000 --terms 500 --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02av_bench.log 2>&1
