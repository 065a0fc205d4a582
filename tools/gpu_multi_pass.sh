#!/bin/bash
# N-GPU pass: the sharded GPU tests, then the sharded headline bench as the driver launches it.
# usage: gpu_multi_pass.sh N TIMEOUT [bench args]
set -u
N=${1:-2}; T=${2:-420}; shift 2
OUT=gpurun_out
mkdir -p $OUT
timeout -s KILL 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > $OUT/r2n${N}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/r2n${N}_pytest.log
timeout $T python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > $OUT/r2n${N}_bench.log 2> $OUT/r2n${N}_bench.err
echo "bench rc=$?"; grep "bench rank 0" $OUT/r2n${N}_bench.err | tail -8; tail -c 600 $OUT/r2n${N}_bench.err; cut -c1-2500 $OUT/r2n${N}_bench.log
