#!/bin/bash
# final tree: the GPU suite as the driver runs it, then once more with every cooperative form forced off and on
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02am_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02am_pytest.log
CB200_COOP=0 CB200_FROBPERS=0 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02am_pytest_coop0.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02am_pytest_coop0.log
CB200_COOP=1 CB200_FROBPERS=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02am_pytest_coop1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02am_pytest_coop1.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02am_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02am_smoke.log
( time python bench.py ) > gpurun_out/r02am_bench1.json 2> gpurun_out/r02am_bench1.err; echo rc=$?
