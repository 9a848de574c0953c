#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "lanes or cooperating or two_lane or frobenius" > gpurun_out/r02af_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02af_pytest.log
tail -3 gpurun_out/r02af_pytest.log
timeout 900 python tools/quick_coop.py > gpurun_out/r02af_quick.log 2>&1
for c in 2 4; do
  echo "COOP=$c frob-calls=2" >> gpurun_out/r02af_bench.log
  CB200_COOP=$c timeout 900 python bench.py --config 2 --frob-calls 2 --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02af_bench.log 2>&1
done
