#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "cooperating or config2 or config4 or hybrid or frobenius or horner or continuation" > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -5 gpurun_out/r02d_pytest.log
for c in 0 1; do
  echo "CB200_COOP=$c" >> gpurun_out/r02d_times.log
  CB200_COOP=$c python tools/bench_configs.py --lean --only 3 >> gpurun_out/r02d_times.log 2>&1
  CB200_COOP=$c python tools/bench_configs.py --lean --only 5 --scale 0.25 >> gpurun_out/r02d_times.log 2>&1
done
cat gpurun_out/r02d_times.log
