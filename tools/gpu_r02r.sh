#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
tail -5 gpurun_out/r02r_pytest.log
for c in 0 2 4; do
  for b in 16384; do
    echo "CB200_COOP=$c batch=$b config2" >> gpurun_out/r02r_frob.log
    CB200_COOP=$c timeout 600 python bench.py --config 2 --batch $b --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02r_frob.log 2>&1
  done
done
echo "config4" >> gpurun_out/r02r_frob.log
timeout 600 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02r_frob.log 2>&1
echo "config1" >> gpurun_out/r02r_frob.log
timeout 600 python bench.py --config 1 --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02r_frob.log 2>&1
timeout 900 python tools/quick_coop.py >> gpurun_out/r02r_quick.log 2>&1
