#!/bin/bash
# shared-operator Fuchs recurrence at 2048 / 4096 bits: one thread per instance against four lanes per instance
set -x
mkdir -p gpurun_out
for L in 64 128; do for b in 256 1024 4096 16384; do for c in 0 1; do
  echo "L=$L batch=$b COOP=$c" >> gpurun_out/r02an.log
  CB200_COOP=$c timeout 300 python tools/quick_time.py $L $b 60 2>&1 | tail -2 | head -1 >> gpurun_out/r02an.log
done; done; done
