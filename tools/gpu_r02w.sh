#!/bin/bash
# ncu --set full with CUDA-source correlation: frobenius1_kernel<128>, the rho = 0 run of configs[2] (first launch)
set -x
mkdir -p gpurun_out
P="python tools/bench_configs.py --lean"
prof() {  # name, kernel regex, skip, args...
  local name=$1 rx=$2 skip=$3; shift 3
  $P "$@" > gpurun_out/${name}_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o /tmp/$name $P "$@" > gpurun_out/${name}_ncu.log 2>&1
  if [ -f /tmp/$name.ncu-rep ]; then
    ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>&1
    ncu -i /tmp/$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
    ncu -i /tmp/$name.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_cuda.csv.gz
  fi
}
export CB200_COOP=0
prof r02w_frob128_rho0 frobenius1 0 --only 3 --terms 12
prof r02w_frob128_dense frobenius1 1 --only 3 --terms 12
