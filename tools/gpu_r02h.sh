#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "fuchs or persistent or full_config or coefficient_ranges" > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
tail -5 gpurun_out/r02h_pytest.log
python tools/bench_configs.py --lean --only 2 > gpurun_out/r02h_sweep.jsonl 2>&1
cat gpurun_out/r02h_sweep.jsonl
P="python tools/bench_configs.py --lean --only 2 --terms 40"
$P > gpurun_out/r02h_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fuchs_table -c 1 -o /tmp/r02h_table $P > gpurun_out/r02h_ncu.log 2>&1
ncu -i /tmp/r02h_table.ncu-rep --page details > gpurun_out/r02h_table_details.txt 2>&1
ncu -i /tmp/r02h_table.ncu-rep --page raw --csv > gpurun_out/r02h_table_raw.csv 2>&1
ncu -i /tmp/r02h_table.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r02h_table_source.csv.gz
