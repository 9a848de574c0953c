#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "integer or config3 or legendre" > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -5 gpurun_out/r02c_pytest.log
python tools/bench_configs.py --lean --only 4 > gpurun_out/r02c_c4.jsonl 2>&1
cat gpurun_out/r02c_c4.jsonl
P="python tools/bench_configs.py --lean --only 4 --terms 200"
$P > gpurun_out/r02c_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fuchs_intop -c 1 -o /tmp/r02c_intop $P > gpurun_out/r02c_ncu.log 2>&1
ncu -i /tmp/r02c_intop.ncu-rep --page details > gpurun_out/r02c_intop_details.txt 2>&1
ncu -i /tmp/r02c_intop.ncu-rep --page raw --csv > gpurun_out/r02c_intop_raw.csv 2>&1
ncu -i /tmp/r02c_intop.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r02c_intop_source.csv.gz
