#!/bin/bash
set -x
mkdir -p gpurun_out
for cfg in "64 65536" "64 262144" "128 65536"; do for c in 0 1; do
  echo "L,batch=$cfg COOP=$c" >> gpurun_out/r02ao.log
  CB200_COOP=$c timeout 300 python tools/quick_time.py $cfg 60 2>&1 | tail -2 | head -1 >> gpurun_out/r02ao.log
done; done
