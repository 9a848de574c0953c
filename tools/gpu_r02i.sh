#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "tma or horner or continuation or config4" > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -5 gpurun_out/r02i_pytest.log
for t in 0 1; do
  echo "CB200_TMA=$t" >> gpurun_out/r02i_times.log
  CB200_TMA=$t timeout 300 python tools/bench_configs.py --lean --only 5 --scale 0.25 >> gpurun_out/r02i_times.log 2>&1
done
cat gpurun_out/r02i_times.log
