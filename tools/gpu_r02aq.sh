#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r02aq_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02aq_pytest.log
CB200_COOP=0 timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r02aq_pytest0.log 2>&1
