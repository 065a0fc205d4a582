#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_bm25.py tests/test_gpu_at_size.py -m gpu -x -q > $OUT/r2c_pytest.log 2>&1
rc=$?
echo "pytest rc=$rc"; tail -4 $OUT/r2c_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2c_bench.log 2> $OUT/r2c_bench.err
echo "bench rc=$?"; tail -c 400 $OUT/r2c_bench.err
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-supplements --no-latency"
ncu --set full --clock-control none --import-source on -k regex:bm25_score -s 3 -c 1 -o $OUT/bm25_score_c5_r02b $BENCH > $OUT/ncu_bm25_r02b.log 2>&1
echo "ncu rc=$?"
