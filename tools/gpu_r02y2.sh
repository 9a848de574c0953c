#!/bin/bash
# round 2, call y2 (final build, after the persistent 4096-bit Frobenius kernel): full GPU suite, bench.py for every config, and the ncu captures behind roofline.traffic
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02y2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02y2_pytest.log
tail -14 gpurun_out/r02y2_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02y2_bench1.json 2> gpurun_out/r02y2_bench1.err; echo "rc=$?"
for c in 2 3 4; do
  python bench.py --config $c --steps 2 --warmup 3 > gpurun_out/r02y2_bench$c.json 2> gpurun_out/r02y2_bench$c.err; echo "cfg $c rc=$?"
done
# launch list of the default bench (kernel shares of the step)
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/r02y2_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02y2_launches1.csv $B > gpurun_out/r02y2_ncu_l1.log 2>&1
# ncu --set full of the dominant kernel of configs 1-3 at their full sizes (traffic per launch); config 4's launch
# takes 8 s, so its capture is the DRAM / duration metrics only (two replay passes instead of forty)
prof() {  # name regex config
  local name=$1 rx=$2 cfg=$3 set=$4
  local C="python bench.py --config $cfg --steps 1 --warmup 1 --no-e2e --no-cpu"
  $C > gpurun_out/${name}_plain.log 2>&1 && \
  ncu $set --clock-control none -k regex:$rx -c 1 -o /tmp/$name $C > gpurun_out/${name}_ncu.log 2>&1
  if [ -f /tmp/$name.ncu-rep ]; then
    ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>&1
    ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>&1
  fi
}
prof r02y2_ncu_cfg1 fuchs_shared_persistent 1 "--set full --import-source on"
prof r02y2_ncu_cfg2 frobenius1 2 "--set full --import-source on"
prof r02y2_ncu_cfg3 fuchs_intop 3 "--set full --import-source on"
prof r02y2_ncu_cfg4 horner 4 "--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__warps_active.avg.per_cycle_active,smsp__issue_active.avg.pct_of_peak_sustained_active"
ncu -i /tmp/r02y2_ncu_cfg2.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r02y2_ncu_cfg2_source.csv.gz
ls -la gpurun_out | grep r02y2
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02y2_smoke.log
timeout 900 python tools/quick_coop.py > gpurun_out/r02y2_quick_coop.log 2>&1
timeout 600 python tools/bench_configs.py --lean --only 2 > gpurun_out/r02y2_pi.log 2>&1
timeout 600 python tools/bench_configs.py --lean --only 9 --continuation --scale 0.03 > gpurun_out/r02y2_cont.log 2>&1
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02y2_ref1.json 2> gpurun_out/r02y2_ref1.err; echo rc=$?
