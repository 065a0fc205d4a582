#!/usr/bin/env python
"""Sweep of the BM25 score-kernel parameters on the C2 workload (1M docs, 1024 queries, top-1000).

For every configuration: K timed batches (a different query batch each), the mean device time of the
prepare / score / select kernels (CUDA events recorded inside the library) and the whole-step time.
The first configuration's results are the comparison baseline: every other configuration must return
the same doc ids (ties aside) and scores within 1e-6 relative.

    python tools/bench_bm25_sweep.py --configs "readout=1;readout=0;range=2048,qpi=4"
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KEYS = {"range": "bm25_range_docs", "readout": "bm25_readout", "qpi": "bm25_queries_per_item", "tau": "bm25_use_tau",
        "candcap": "bm25_cand_cap", "init": "bm25_tau_init", "accum": "bm25_accum"}
DEFAULTS = {"range": 0, "readout": 1, "qpi": 0, "tau": 1, "candcap": 0, "init": 1, "accum": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="readout=1;readout=0")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--docs", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--top-k", type=int, default=1000)
    ap.add_argument("--always", type=float, default=0.0, help="fraction of docs holding the always-term appended to every query (C5: 0.95)")
    a = ap.parse_args()

    import torch
    import mse_b200  # noqa: F401
    from mse_b200 import _native, synthetic
    dev = torch.device("cuda", 0)
    c = synthetic.make_bm25_corpus(a.docs, vocab=200_000, seed=1234, device=dev, always_frac=a.always)
    nat = _native.NativeIndex(0)
    nat.bm25_load(c.term_off, c.post_doc, c.post_tf, c.doc_len, c.idf, c.avgdl)
    if a.always > 0:
        nat.set_option("bm25_class_term", c.always_term)      # as bench.py / the pipeline do for the appended term
    df = torch.diff(c.term_off).cpu().numpy()
    n = a.steps + a.warmup
    batches, posts = [], []
    for i in range(n):
        q_off, q_term, q_tf = synthetic.make_bm25_queries(c, a.batch, seed=1235 + i, add_always=a.always > 0)
        batches.append(tuple(torch.from_numpy(x).to(dev) for x in (q_off, q_term, q_tf)))
        posts.append(int(df[q_term].sum()))
    alg = 12.0 * float(np.mean(posts[a.warmup:])) + 8.0 * a.top_k * a.batch
    out = (torch.empty((a.batch, a.top_k), dtype=torch.int32, device=dev), torch.empty((a.batch, a.top_k), dtype=torch.float32, device=dev),
           torch.empty((a.batch,), dtype=torch.int32, device=dev))
    base = None
    for spec in a.configs.split(";"):
        cfg = dict(DEFAULTS)
        for kv in filter(None, spec.split(",")):
            k, v = kv.split("=")
            cfg[k.strip()] = int(v)
        min_score = cfg.pop("ms10", 0) / 10.0
        for k, v in cfg.items():
            nat.set_option(KEYS[k], v)
        rec = {"config": spec}
        try:
            for i in range(a.warmup):
                nat.bm25_search(*batches[i], a.top_k, min_score, out=out)
            torch.cuda.synchronize()
            nat.set_option("reset_timers", 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in range(a.steps):
                nat.bm25_search(*batches[a.warmup + s], a.top_k, min_score, out=out)
            e1.record()
            torch.cuda.synchronize()
            sc, scn = nat.kernel_time("bm25_score")
            se, _ = nat.kernel_time("topk_select")
            pr, _ = nat.kernel_time("bm25_prepare")
            st = nat.bm25_stats()
            rec.update(step_ms=e0.elapsed_time(e1) / a.steps, score_ms=sc / max(1, scn), prepare_ms=pr / max(1, scn),
                       select_ms=se / max(1, scn), score_GBps=alg / (sc / max(1, scn) * 1e-3) / 1e9,
                       emitted_per_query=st.get("emitted", 0) / a.batch, reruns=st.get("rerun_queries", 0),
                       ctas=st.get("ctas"), exact_tasks=st.get("exact_mode_tasks"))
            # comparison on the LAST timed batch
            ids, scores, cnt = (t.clone().cpu().numpy() for t in out)
            if base is None:
                base = (ids, scores, cnt)
                rec["vs_first"] = "baseline"
            else:
                bi, bs, bc = base
                same_cnt = bool(np.array_equal(cnt, bc))
                rel = float(np.max(np.abs(scores - bs) / np.maximum(np.abs(bs), 1e-6)))
                id_match = float(np.mean(ids == bi))
                rec["vs_first"] = {"counts_equal": same_cnt, "max_rel_score_diff": rel, "id_match_frac": id_match}
        except Exception as e:  # noqa: BLE001
            rec["error"] = repr(e)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
