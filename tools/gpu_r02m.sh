#!/bin/bash
# ncu --set full of the cooperative Frobenius kernels (2 and 4 lanes per instance) on a short configs[2] run
set -x
mkdir -p gpurun_out
P="python tools/bench_configs.py --lean"
prof() {  # name, kernel regex, args...
  local name=$1 rx=$2; shift 2
  $P "$@" > gpurun_out/${name}_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -o /tmp/$name $P "$@" > gpurun_out/${name}_ncu.log 2>&1
  if [ -f /tmp/$name.ncu-rep ]; then
    ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>&1
    ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>&1
    ncu -i /tmp/$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
  fi
}
export CB200_COOP=2
prof r02m_frob128_2lane frobenius1_coop --only 3 --terms 12
export CB200_COOP=4
prof r02m_frob128_4lane frobenius1_coop --only 3 --terms 12
ls -la gpurun_out
