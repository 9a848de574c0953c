#!/bin/bash
# ncu --set full with CUDA-source correlation of the headline kernel (configs[1], short series)
set -x
mkdir -p gpurun_out
P="python bench.py --config 1 --terms 200 --steps 1 --warmup 1 --no-e2e --no-cpu"
name=r02y_cfg1
$P > gpurun_out/${name}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fuchs_shared_persistent -c 1 -o /tmp/$name $P > gpurun_out/${name}_ncu.log 2>&1
if [ -f /tmp/$name.ncu-rep ]; then
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>&1
  ncu -i /tmp/$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
  ncu -i /tmp/$name.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_cuda.csv.gz
fi
