#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02au_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02au_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02au_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02au_smoke.log
for cfg in "64 16384" "128 16384"; do
  echo "L,batch=$cfg" >> gpurun_out/r02au.log
  timeout 300 python tools/quick_time.py $cfg 60 2>&1 | tail -2 | head -1 >> gpurun_out/r02au.log
done
timeout 600 python tools/bench_configs.py --lean --only 9 --continuation --scale 0.03 >> gpurun_out/r02au_cont.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02au_bench1.json 2> gpurun_out/r02au_bench1.err
