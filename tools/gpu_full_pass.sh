#!/bin/bash
# the whole -m gpu suite, then the default bench run (as the driver launches it) and the reference arm
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/r2f_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 $OUT/r2f_pytest.log
timeout -s KILL 900 python bench.py > $OUT/r2f_bench.log 2> $OUT/r2f_bench.err
echo "bench rc=$?"; tail -c 400 $OUT/r2f_bench.err
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r2f_ref.log 2> $OUT/r2f_ref.err
echo "ref rc=$?"; tail -c 300 $OUT/r2f_ref.err; cut -c1-600 $OUT/r2f_ref.log
