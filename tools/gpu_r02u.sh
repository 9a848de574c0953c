#!/bin/bash
set -x
mkdir -p gpurun_out
for lib in libcascade_b200 libcb_h320 libcb_h384; do
  export CB200_LIB=$PWD/cascade_b200/$lib.so
  for tma in 0 1; do
    echo "$lib TMA=$tma" >> gpurun_out/r02u_bench.log
    CB200_TMA=$tma timeout 600 python bench.py --config 4 --terms 500 --steps 2 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r02u_bench.log 2>&1
  done
  timeout 600 python -m pytest tests -m gpu -x -q -k "horner" > gpurun_out/r02u_pytest_$lib.log 2>&1; echo "rc=$?" >> gpurun_out/r02u_pytest_$lib.log
done
