#!/usr/bin/env python
"""End-to-end hybrid query at the BASELINE.json configs[4] size on ONE B200:
BM25 candidates over N docs (default 10M; ~1.9G postings) -> gathered chunk rerank (5 chunks/doc = 50M x 768 bf16
chunks, 76.8 GB) + score fusion -> top-100, for batch 1 (latency) and batch 4096 (throughput).
Everything is device-resident; one JSON line per batch size.

    python tools/bench_hybrid_c5.py [--docs 10000000] [--batches 1,4096] [--steps 5]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--vocab", type=int, default=200_000)
    ap.add_argument("--chunks-per-doc", type=int, default=5)
    ap.add_argument("--batches", default="1,4096")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--top-k", type=int, default=1000)
    ap.add_argument("--max-out", type=int, default=100)
    a = ap.parse_args()

    import torch
    import mse_b200  # noqa: F401
    from mse_b200 import _native, synthetic
    dev = torch.device("cuda", 0)
    t0 = time.time()
    c = synthetic.make_bm25_corpus(a.docs, vocab=a.vocab, seed=1234, device=dev)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    nat = _native.NativeIndex(0)
    t0 = time.time()
    nat.bm25_load(c.term_off, c.post_doc, c.post_tf, c.doc_len, c.idf, c.avgdl)
    torch.cuda.synchronize()
    t_load = time.time() - t0
    n_postings = int(c.n_postings)
    df = torch.diff(c.term_off).cpu().numpy()
    # queries are generated from the corpus statistics before the posting arrays are dropped
    batches = {}
    for B in [int(x) for x in a.batches.split(",")]:
        batches[B] = [tuple(torch.from_numpy(x).to(dev) for x in synthetic.make_bm25_queries(c, B, seed=4321 + 17 * i + B))
                      for i in range(a.steps + 2)]
    posts = {B: float(np.mean([df[b[1].cpu().numpy()].sum() for b in bs])) for B, bs in batches.items()}
    c.post_doc = c.post_tf = None
    del c
    torch.cuda.empty_cache()
    t0 = time.time()
    d = synthetic.make_dense_corpus(a.docs, seed=1234, device=dev, dtype=torch.bfloat16, chunks_per_doc=a.chunks_per_doc)
    nat.dense_load(d.emb, d.doc_chunk_off, borrow=True)          # 76.8 GB: scanned in place, not copied
    torch.cuda.synchronize()
    t_dense = time.time() - t0
    print(f"# corpus: {a.docs} docs, {n_postings} postings (gen {t_gen:.1f}s, load {t_load:.1f}s), "
          f"{a.docs * a.chunks_per_doc} chunks (gen+load {t_dense:.1f}s), "
          f"{torch.cuda.memory_allocated() / 2**30:.1f} GiB allocated by torch", file=sys.stderr)

    for B, bs in batches.items():
        qv = torch.from_numpy(synthetic.make_query_vectors(B, seed=99)).to(dev)
        cand_off = (torch.arange(B + 1, device=dev, dtype=torch.int32) * a.top_k).contiguous()

        def step(i):
            q_off, q_term, q_tf = bs[i]
            doc, score, count = nat.bm25_search(q_off, q_term, q_tf, a.top_k, 0.0)
            return nat.rerank(cand_off, doc.view(-1), score.view(-1), qv, None, 0.15, 10, a.max_out), count

        for i in range(2):
            step(i)
        torch.cuda.synchronize()
        nat.set_option("reset_timers", 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            out, count = step(2 + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        sc, n = nat.kernel_time("bm25_score")
        se, _ = nat.kernel_time("topk_select")
        pr, _ = nat.kernel_time("bm25_prepare")
        rr, nr = nat.kernel_time("rerank")
        st = nat.bm25_stats()
        rows = float(out[5].float().mean().item())
        print(json.dumps({
            "workload": f"C5 hybrid: BM25 top-{a.top_k} over {a.docs} docs ({n_postings} postings) -> rerank <=10 of "
                        f"{a.chunks_per_doc} chunks/doc (768-d bf16, {a.docs * a.chunks_per_doc} chunks) -> top-{a.max_out}",
            "batch": B, "ms_per_batch": ms, "hybrid_queries_per_s": B / (ms / 1e3),
            "bm25_score_ms": sc / max(n, 1), "bm25_prepare_ms": pr / max(n, 1), "select_ms": se / max(n, 1),
            "rerank_ms": rr / max(nr, 1), "postings_per_query": posts[B] / B,
            "bm25_score_GBps_algorithmic": (12.0 * posts[B] + 8.0 * a.top_k * B) / (sc / max(n, 1) * 1e-3) / 1e9,
            "rerank_GBps_algorithmic": B * (2.0 * 768 * rows + 12.0 * a.top_k + 8.0 * a.max_out) / (rr / max(nr, 1) * 1e-3) / 1e9,
            "candidates_emitted_per_query": st["emitted"] / B, "rerun_queries": st["rerun_queries"],
            "rerank_rows_per_query": rows}), flush=True)


if __name__ == "__main__":
    main()
