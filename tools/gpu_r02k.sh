#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cooperating or continuation or monodromy or fuchs" > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log
tail -5 gpurun_out/r02k_pytest.log
for c in 0 1; do
  echo "CB200_COOP=$c" >> gpurun_out/r02k_cont.log
  CB200_COOP=$c timeout 600 python tools/bench_configs.py --lean --only 9 --continuation --scale 0.03 >> gpurun_out/r02k_cont.log 2>&1
done
timeout 600 python tools/bench_configs.py --lean --only 9 --continuation --scale 0.03 >> gpurun_out/r02k_cont.log 2>&1
cat gpurun_out/r02k_cont.log
python bench.py --steps 2 --warmup 3 --no-cpu --e2e-tiles 4 > gpurun_out/r02k_bench1_tiles4.json 2> gpurun_out/r02k_bench1_tiles4.err; echo rc=$?
python bench.py --steps 2 --warmup 3 --no-cpu --e2e-tiles 2 > gpurun_out/r02k_bench1_tiles2.json 2> gpurun_out/r02k_bench1_tiles2.err; echo rc=$?
