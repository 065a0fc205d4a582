#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -x -q > $OUT/r2d_pytest.log 2>&1
rc=$?
echo "dense pytest rc=$rc"; tail -15 $OUT/r2d_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests/test_gpu_at_size.py -m gpu -x -q -k c3 > $OUT/r2d_pytest2.log 2>&1
echo "c3 rc=$?"; tail -5 $OUT/r2d_pytest2.log
timeout 300 python tools/bench_dense.py --batches 1,8,64,128,256 --steps 5 > $OUT/r2d_dense.log 2>&1
echo "bench_dense rc=$?"; tail -8 $OUT/r2d_dense.log | cut -c1-420
timeout 300 python tools/bench_dense.py --batches 128,256 --steps 5 --gemm-debug 1 > $OUT/r2d_dense_dbg.log 2>&1
echo "debug=1"; tail -2 $OUT/r2d_dense_dbg.log | cut -c60-260
