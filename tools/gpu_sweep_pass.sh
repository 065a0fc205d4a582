#!/bin/bash
# score-kernel variants on the C5 BM25 stage (10 M docs, always-term, batch 4096) and optionally on C2; every
# configuration is compared with the first one
set -u
OUT=gpurun_out
mkdir -p $OUT
CFG=${1:-"var=0;var=7;var=0"}
timeout -s KILL 600 python tools/bench_bm25_sweep.py --docs 10000000 --batch 4096 --always 0.95 --steps 6 --warmup 3 --configs "$CFG" > $OUT/sweep_c5.log 2> $OUT/sweep_c5.err
echo "c5 rc=$?"; tail -c 600 $OUT/sweep_c5.err
if [ "${2:-}" = "c2" ]; then
timeout -s KILL 300 python tools/bench_bm25_sweep.py --steps 20 --warmup 3 --configs "$CFG" > $OUT/sweep_c2.log 2> $OUT/sweep_c2.err
echo "c2 rc=$?"; tail -c 600 $OUT/sweep_c2.err
fi
python - <<'PY'
import json, glob
for f in ("gpurun_out/sweep_c5.log", "gpurun_out/sweep_c2.log"):
    try: lines = open(f).read().splitlines()
    except OSError: continue
    print(f)
    for l in lines:
        if l.startswith("{"):
            j = json.loads(l)
            v = j.get("vs_first")
            print(" ", j["config"], "score", round(j.get("score_ms", -1), 3), "step", round(j.get("step_ms", -1), 3), "emit", round(j.get("emitted_per_query", -1)), "exact", j.get("exact_tasks"),
                  v if isinstance(v, str) else (v and (v["counts_equal"], v["max_rel_score_diff"], v["id_match_frac"])), j.get("error", ""))
PY
