#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "horner or cooperating or continuation or evaluate" > gpurun_out/r02ai_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ai_pytest.log
tail -3 gpurun_out/r02ai_pytest.log
timeout 900 python tools/quick_coop.py > gpurun_out/r02ai_quick.log 2>&1
