#!/bin/bash
# round 2, first GPU call: ncu --set full of the three kernels that had no profile in round 1.
# The .ncu-rep files are converted to text on the box (gpurun_out is capped at 64 MiB) and removed.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt
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
prof r02a_frob128 frobenius1 --only 3 --terms 12
prof r02a_intop32 fuchs_intop --only 4 --terms 200
prof r02a_table32 fuchs_table --only 2 --terms 40
python tools/bench_configs.py > gpurun_out/r02a_configs.jsonl 2> gpurun_out/r02a_configs.err
ls -la gpurun_out
