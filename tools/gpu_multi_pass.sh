#!/bin/bash
# N-GPU pass: the sharded headline bench as the driver launches it.  usage: gpu_round2_n2.sh N TIMEOUT [bench args]
set -u
N=${1:-2}; T=${2:-420}; shift 2
OUT=gpurun_out
mkdir -p $OUT
timeout $T python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > $OUT/r2n${N}_bench.log 2> $OUT/r2n${N}_bench.err
echo "bench rc=$?"; grep "bench rank 0" $OUT/r2n${N}_bench.err | tail -8; tail -c 600 $OUT/r2n${N}_bench.err; cut -c1-2500 $OUT/r2n${N}_bench.log
