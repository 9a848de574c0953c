#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fuchs or continuation or cooperating or monodromy" > gpurun_out/r02at_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02at_pytest.log
for cfg in "64 256" "64 16384" "64 262144" "128 256" "128 16384" "128 65536"; do
  echo "L,batch=$cfg sync" >> gpurun_out/r02at.log
  timeout 300 python tools/quick_time.py $cfg 60 2>&1 | tail -2 | head -1 >> gpurun_out/r02at.log
done
timeout 600 python tools/bench_configs.py --lean --only 9 --continuation --scale 0.03 >> gpurun_out/r02at_cont.log 2>&1
