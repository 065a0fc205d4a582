#!/bin/bash
# first GPU pass of round 2: the whole -m gpu suite, then a short full bench run
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv > $OUT/r2a_gpu.txt 2>&1
free -g >> $OUT/r2a_gpu.txt; nproc >> $OUT/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/r2a_pytest.log 2>&1
echo "pytest rc=$?"
tail -5 $OUT/r2a_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r2a_bench.log 2> $OUT/r2a_bench.err
echo "bench rc=$?"
tail -c 600 $OUT/r2a_bench.err
