#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout -s KILL 150 python -m pytest tests/test_gpu_dense.py -m gpu -x -q > $OUT/r2d_pytest.log 2>&1
rc=$?
echo "dense pytest rc=$rc"; tail -15 $OUT/r2d_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout -s KILL 200 python tools/bench_dense.py --batches 64,128,256 --steps 5 > $OUT/r2d_dense.log 2>&1
echo "bench_dense rc=$?"; tail -3 $OUT/r2d_dense.log | cut -c1-330
timeout -s KILL 200 python tools/bench_dense.py --batches 256 --steps 5 --gemm-debug 1 > $OUT/r2d_dense_dbg.log 2>&1
echo "debug=1"; tail -1 $OUT/r2d_dense_dbg.log | cut -c60-260
timeout -s KILL 300 python -m pytest tests/test_gpu_at_size.py -m gpu -x -q -k c3 > $OUT/r2d_pytest2.log 2>&1
echo "c3 rc=$?"; tail -5 $OUT/r2d_pytest2.log
