#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lanes or cooperating or two_lane" > gpurun_out/r02ab_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ab_pytest.log
timeout 900 python tools/quick_coop.py > gpurun_out/r02ab_quick.log 2>&1
timeout 600 python tools/bench_configs.py --lean --only 9 --continuation --scale 0.03 > gpurun_out/r02ab_cont.log 2>&1
