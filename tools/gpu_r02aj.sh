#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02aj_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02aj_pytest.log
tail -3 gpurun_out/r02aj_pytest.log
timeout 900 python tools/quick_coop.py > gpurun_out/r02aj_quick.log 2>&1
