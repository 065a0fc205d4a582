#!/bin/bash
# BM25 kernel tests, then the score-kernel sweep (C5 BM25 stage, optionally C2)
cd /root/repo
timeout -s KILL 600 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/u16_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/u16_pytest.log
bash tools/gpu_sweep_pass.sh "${1:-accum=32;accum=0}" ${2:-}
