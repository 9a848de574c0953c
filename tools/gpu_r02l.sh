#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "two_lanes or cooperating" > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest.log
tail -5 gpurun_out/r02l_pytest.log
for c in 0 2 4; do
  for b in 16384 2048 512; do
    echo "CB200_COOP=$c batch=$b config2" >> gpurun_out/r02l_frob.log
    CB200_COOP=$c timeout 600 python bench.py --config 2 --batch $b --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02l_frob.log 2>&1
  done
done
cat gpurun_out/r02l_frob.log | cut -c1-400
