#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "frobenius or config2" > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.log
timeout 900 python bench.py --config 2 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02z_bench2.json 2> gpurun_out/r02z_bench2.err
timeout 600 python tools/bench_configs.py --lean --only 3 > gpurun_out/r02z_cfg.log 2>&1
