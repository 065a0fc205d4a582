#!/bin/bash
# BM25 / rerank pass: kernel tests, headline bench without the supplements
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_api.py tests/test_gpu_dense.py tests/test_gpu_pipeline.py -m gpu -x -q > $OUT/r2e_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -3 $OUT/r2e_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout -s KILL 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-supplements --no-latency > $OUT/r2e_bench.log 2> $OUT/r2e_bench.err
echo "bench rc=$?"; tail -c 300 $OUT/r2e_bench.err
python - <<'PY'
import json
j=json.loads([x for x in open('gpurun_out/r2e_bench.log') if x.startswith('{')][-1])
print(j['value'], j['ms_per_step'], j['roofline']['frac'], j['breakdown']['score_ms'], j['breakdown']['prepare_ms'], j['breakdown']['select_ms'], j['breakdown']['rerank_ms'], j['rooflines'][1]['frac'])
print(json.dumps(j['parity'])[:400])
PY
