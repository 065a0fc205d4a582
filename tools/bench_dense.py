#!/usr/bin/env python
"""Dense exhaustive scan micro-benchmark (BASELINE.json configs[2]: 10M x 768 bf16 chunks, 2M docs,
per-doc max-pool + top-1000).  Prints one JSON line per batch size with the scan kernel's HBM GB/s."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mse_b200  # noqa
from mse_b200 import _native, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=10_000_000)
ap.add_argument("--chunks-per-doc", type=int, default=5)
ap.add_argument("--batches", default="1,2,4,8")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--top-k", type=int, default=1000)
ap.add_argument("--ctas-per-sm", type=int, default=0)
ap.add_argument("--gemm-debug", type=int, default=0)
ap.add_argument("--pair-mode", type=int, default=0, help="dense_gemm_pair_mode (1: multicast pair, 2: cta_group::2 pair)")
ap.add_argument("--slab-rows", action="store_true", help="the slab-seeded row generator bench.py's C3 supplement uses")
a = ap.parse_args()
dev = torch.device("cuda:0")
n_docs = a.chunks // a.chunks_per_doc
t0 = time.time()
nat = _native.NativeIndex(0)
if a.slab_rows:
    emb = synthetic.dense_rows(0, a.chunks, 1234, device=dev, dtype=torch.bfloat16)
    nat.dense_load(emb, torch.arange(n_docs + 1, dtype=torch.int64, device=dev) * a.chunks_per_doc, borrow=True)
else:
    d = synthetic.make_dense_corpus(n_docs, seed=1234, device=dev, dtype=torch.bfloat16, chunks_per_doc=a.chunks_per_doc)
    nat.dense_load(d.emb, d.doc_chunk_off)
    del d.emb
torch.cuda.empty_cache()
if a.pair_mode:
    nat.set_option("dense_gemm_pair_mode", a.pair_mode)
if a.gemm_debug:
    nat.set_option("dense_gemm_debug", a.gemm_debug)
if a.ctas_per_sm:
    nat.set_option("dense_scan_ctas_per_sm", a.ctas_per_sm)
print(f"# corpus ready in {time.time()-t0:.1f}s", file=sys.stderr)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
for B in [int(x) for x in a.batches.split(",")]:
    q = torch.from_numpy(synthetic.make_query_vectors(B, seed=77, normalize=True)).to(dev)
    for _ in range(3):
        nat.dense_scan(q, a.top_k)
    torch.cuda.synchronize()
    nat.set_option("reset_timers", 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = nat.dense_scan(q, a.top_k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    scan_ms, n = nat.kernel_time("dense_scan"); sel_ms, _ = nat.kernel_time("topk_select")
    scan_ms /= max(n, 1); sel_ms /= max(n, 1)
    passes = (B + 3) // 4 if B >= 4 else 1
    alg = 2.0 * 768 * a.chunks + 8.0 * (n_docs + 1) + 4.0 * 768 * B + 8.0 * a.top_k * B
    print(json.dumps({"workload": f"dense scan {a.chunks} x 768 bf16, {n_docs} docs, B={B}, top-{a.top_k}", "ms_per_batch": ms,
                      "queries_per_s": B / (ms / 1e3), "scan_ms": scan_ms, "select_ms": sel_ms,
                      "scan_GBps_algorithmic": alg / (scan_ms * 1e-3) / 1e9, "frac_of_measured_hbm": alg / (scan_ms * 1e-3) / 1e9 / peak,
                      "matrix_passes": passes}))
