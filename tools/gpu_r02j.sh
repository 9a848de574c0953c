#!/bin/bash
set -x
mkdir -p gpurun_out
free -g > gpurun_out/r02j_free.txt; nproc >> gpurun_out/r02j_free.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
tail -4 gpurun_out/r02j_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02j_smoke.log
( time python bench.py ) > gpurun_out/r02j_bench1.json 2> gpurun_out/r02j_bench1.err; echo rc=$?
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02j_ref1.json 2> gpurun_out/r02j_ref1.err; echo rc=$?
python bench.py --config 3 --steps 2 --warmup 3 > gpurun_out/r02j_bench3.json 2> gpurun_out/r02j_bench3.err
tail -3 gpurun_out/r02j_bench1.err gpurun_out/r02j_ref1.err
