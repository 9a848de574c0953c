#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "two_lanes or cooperating or frobenius or elementwise" > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p_pytest.log
tail -5 gpurun_out/r02p_pytest.log
for c in 2 4; do
  for b in 16384 2048; do
    echo "CB200_COOP=$c batch=$b config2" >> gpurun_out/r02p_frob.log
    CB200_COOP=$c timeout 600 python bench.py --config 2 --batch $b --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02p_frob.log 2>&1
  done
done
timeout 900 python tools/quick_coop.py >> gpurun_out/r02p_quick.log 2>&1
