#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02v_pytest.log
tail -5 gpurun_out/r02v_pytest.log
for c in 0 4; do
  echo "CB200_COOP=$c config2" >> gpurun_out/r02v_bench.log
  CB200_COOP=$c timeout 900 python bench.py --config 2 --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02v_bench.log 2>&1
done
timeout 600 python tools/bench_configs.py --lean --only 3 >> gpurun_out/r02v_cfg.log 2>&1
timeout 600 python tools/bench_configs.py --lean --only 9 --continuation --scale 0.03 >> gpurun_out/r02v_cfg.log 2>&1
