#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "frobenius or config2 or general or compat" > gpurun_out/r02ak_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ak_pytest.log
tail -3 gpurun_out/r02ak_pytest.log
for c in 1 2; do
  echo "frob-calls=$c" >> gpurun_out/r02ak_bench.log
  timeout 900 python bench.py --config 2 --frob-calls $c --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02ak_bench.log 2>&1
done
