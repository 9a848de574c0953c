#!/bin/bash
# last call of the round: the tree as committed
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02as_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02as_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02as_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02as_smoke.log
( time python bench.py --impl reference ) > gpurun_out/r02as_ref1.json 2> gpurun_out/r02as_ref1.err
( time python bench.py ) > gpurun_out/r02as_bench1.json 2> gpurun_out/r02as_bench1.err
for c in 0 1; do
echo "COOP=$c" >> gpurun_out/r02as_cont.log
CB200_COOP=$c timeout 600 python tools/bench_configs.py --lean --only 9 --continuation --continuation-batch 16384 --scale 0.03 >> gpurun_out/r02as_cont.log 2>&1
done
