#!/bin/bash
set -u
set -o pipefail
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/r2b_pytest.log 2>&1
rc=$?
echo "pytest rc=$rc"; tail -4 $OUT/r2b_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r2b_bench.log 2> $OUT/r2b_bench.err
echo "bench rc=$?"; tail -c 400 $OUT/r2b_bench.err
bash tools/profile_round2.sh r02a
