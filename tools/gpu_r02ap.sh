#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02ap_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ap_pytest.log
tail -3 gpurun_out/r02ap_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ap_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02ap_smoke.log
timeout 900 python tools/bench_configs.py --only 5 --continuation > gpurun_out/r02ap_cfg5.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02ap_bench1.json 2> gpurun_out/r02ap_bench1.err
