#!/bin/bash
set -x
mkdir -p gpurun_out
for cfg in "128 256" "128 4096" "128 16384" "128 65536"; do
  echo "L,batch=$cfg TPB192" >> gpurun_out/r02ar.log
  timeout 300 python tools/quick_time.py $cfg 60 2>&1 | tail -2 | head -1 >> gpurun_out/r02ar.log
done
timeout 900 python -m pytest tests -m gpu -x -q -k "fuchs or continuation or cooperating or monodromy" > gpurun_out/r02ar_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ar_pytest.log
