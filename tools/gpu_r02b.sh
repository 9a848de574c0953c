#!/bin/bash
# round 2, call b: the whole GPU test-suite, then bench.py for every config (short runs)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -30 gpurun_out/r02b_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02b_bench1.json 2> gpurun_out/r02b_bench1.err; echo "rc=$?"
for c in 2 3 4; do
  python bench.py --config $c --steps 2 --warmup 3 > gpurun_out/r02b_bench$c.json 2> gpurun_out/r02b_bench$c.err; echo "cfg $c rc=$?"
done
python bench.py --impl reference --steps 1 --warmup 0 --config 3 > gpurun_out/r02b_ref3.json 2> gpurun_out/r02b_ref3.err
tail -c 1500 gpurun_out/r02b_bench*.json
tail -5 gpurun_out/r02b_bench*.err
