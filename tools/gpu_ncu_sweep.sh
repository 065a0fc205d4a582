#!/bin/bash
# one `ncu --set full` capture of the BM25 score kernel on the C5 BM25 stage; usage: gpu_ncu_sweep.sh TAG [kernel regex] [configs]
set -u
TAG=${1:-x}; K=${2:-bm25_score}; CFG=${3:-accum=0}
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/bm25_$TAG python tools/bench_bm25_sweep.py --docs 10000000 --batch 4096 --always 0.95 --steps 2 --warmup 2 --configs "$CFG" > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$TAG.log
