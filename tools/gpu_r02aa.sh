#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02aa_pytest.log
tail -3 gpurun_out/r02aa_pytest.log
for cfg in 1 2 3 4; do
  echo "config$cfg" >> gpurun_out/r02aa_bench.log
  timeout 900 python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02aa_bench.log 2>&1
done
timeout 600 python tools/bench_configs.py --lean --only 2 > gpurun_out/r02aa_pi.log 2>&1
timeout 600 python tools/bench_configs.py --lean --only 9 --continuation --scale 0.03 > gpurun_out/r02aa_cont.log 2>&1
timeout 900 python tools/quick_coop.py > gpurun_out/r02aa_quick.log 2>&1
