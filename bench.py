#!/usr/bin/env python
"""bench.py — headline benchmark of the retrieval hot path.

Workload (BASELINE.json configs[1]): BM25-only, 1M synthetic docs, Zipf(1.0) vocabulary of 200k,
~256 tokens/doc (~192M postings), batches of 1024 four-term queries, top-1000, on one B200.
A "step" = one batch of queries through prepare -> score -> select.  On N > 1 GPUs (weak scaling: every GPU
traverses the same number of postings per step as the single GPU does, the job processes N x 1024 queries):
  * `--layout replicated` (what `auto` picks when the index fits one GPU, as the 1.5 GB C2 index does): the
    QUERIES are the sharded unit — every rank holds the whole index and answers its own 1024-query batch, no
    data-path collective;
  * `--layout doc-sharded` (what `auto` picks for an index that does not fit): the corpus is sharded by doc range
    (global idf / avgdl), each rank scores the whole replicated N x 1024 batch against its shard, keeps
    m = 2k/N + 32 entries per query, the lists travel to the rank that owns the query (NCCL all-to-all over
    NVLink) and are merged there to the exact global top-1000 (a step whose cut could have hidden a result is
    repeated with full lists).  `--exchange allgather` runs all-gather + merge-everywhere instead (batch 1024 at
    every N: strong scaling).
With the replicated layout the doc-sharded exchange is still measured on the same GPUs for a few steps and
reported as the `doc_sharded` supplement of the JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = queries/s with inputs resident in HBM; `e2e` = the same
through the C ABI with HOST buffers (H2D of the query CSR and D2H of the results inside the timed
region); `roofline` describes the dominant kernel (bm25_score_kernel) with algorithmic bytes
12*P_q + 8*k per query; `cpu_baseline` = the oracle port of the reference's Python
scoring loop (indexer/bm25_indexer.py:435-485) timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DOCS = 1_000_000
VOCAB = 200_000
BATCH = 1024
TOP_K = 1000
SEED = 1234
METRIC = "bm25_queries_per_sec"
UNIT = "queries/s"
WORKLOAD = "C2 BM25-only: 1M synthetic docs, Zipf(1.0) vocab 200k, batch 1024 x 4-term queries, top-1000"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index: int, period: float = 0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._halt = threading.Event()
        self.active = threading.Event()
        self.err = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self._halt.is_set():
                if self.active.is_set():
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    r = int(get_reasons(h))
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def stop(self):
        self._halt.set()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": 0, "error": self.err}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def gen_corpus(device):
    from mse_b200 import synthetic
    return synthetic.make_bm25_corpus(N_DOCS, vocab=VOCAB, seed=SEED, device=device)


def shard_corpus(c, rank, world):
    """Contiguous doc-index range balanced by postings; postings re-based to local doc indices."""
    import torch
    if world == 1:
        return c.term_off, c.post_doc, c.post_tf, c.doc_len, 0
    per_doc = torch.bincount(c.post_doc.long(), minlength=c.n_docs)
    cum = torch.cumsum(per_doc, 0)
    total = int(cum[-1].item())
    targets = torch.tensor([total * r // world for r in range(1, world)], device=cum.device)
    cuts = [0] + (torch.searchsorted(cum, targets) + 1).tolist() + [c.n_docs]
    lo, hi = int(cuts[rank]), int(cuts[rank + 1])
    keep = (c.post_doc >= lo) & (c.post_doc < hi)
    df = torch.diff(c.term_off)
    term_of = torch.repeat_interleave(torch.arange(c.n_terms, device=df.device), df)
    ndf = torch.bincount(term_of[keep], minlength=c.n_terms)
    term_off = torch.zeros(c.n_terms + 1, dtype=torch.int64, device=df.device)
    term_off[1:] = torch.cumsum(ndf, 0)
    return term_off, (c.post_doc[keep] - lo).contiguous(), c.post_tf[keep].contiguous(), c.doc_len[lo:hi].contiguous(), lo


def dense_scan_supplement(nat, dev, peak, n_chunks=10_000_000, chunks_per_doc=5, top_k=1000, steps=5):
    """C3: 10M x 768 bf16 chunks, 2M docs, per-doc max-pool + top-1000.  B=1/2 use the GEMV kernel, B>=8 the
    tcgen05 GEMM kernel.  Reports the scan kernel's algorithmic HBM GB/s and, for the GEMM, bf16 TFLOP/s."""
    import torch
    from mse_b200 import synthetic
    n_docs = n_chunks // chunks_per_doc
    d = synthetic.make_dense_corpus(n_docs, seed=SEED, device=dev, dtype=torch.bfloat16, chunks_per_doc=chunks_per_doc)
    nat.dense_load(d.emb, d.doc_chunk_off)
    del d
    torch.cuda.empty_cache()
    out = []
    for B in (1, 8, 64, 256):
        q = torch.from_numpy(synthetic.make_query_vectors(B, seed=77, normalize=True)).to(dev)
        for _ in range(3):
            nat.dense_scan(q, top_k)
        torch.cuda.synchronize()
        nat.set_option("reset_timers", 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            nat.dense_scan(q, top_k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        scan_ms, n = nat.kernel_time("dense_scan")
        scan_ms /= max(n, 1)
        alg = 2.0 * 768 * n_chunks + 8.0 * (n_docs + 1) + 4.0 * 768 * B + 8.0 * top_k * B
        out.append({"batch": B, "kernel": "dense_gemm_kernel (tcgen05)" if B >= 8 else "dense_scan_kernel (GEMV)",
                    "ms_per_batch": ms, "queries_per_s": B / (ms / 1e3), "scan_kernel_ms": scan_ms,
                    "hbm_GBps_algorithmic": alg / (scan_ms * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (scan_ms * 1e-3) / 1e9 / peak,
                    "bf16_TFLOPs": 2.0 * 768 * n_chunks * B / (scan_ms * 1e-3) / 1e12})
    return {"workload": f"C3 dense exhaustive scan: {n_chunks} x 768 bf16 chunks, {n_docs} docs, per-doc max-pool, top-{top_k}",
            "results": out}


def hybrid_supplement(nat, dev, dev_batches, peak, steps=20, chunks_per_doc=5, max_out=100):
    """End-to-end hybrid query (BASELINE.json configs[4] shape at the C2 corpus size): BM25 top-1000 over the
    1M-doc index -> gathered rerank of <= 10 chunks/doc (768-d bf16) with min-max fusion -> top-100, batches of
    1024 queries, everything device-resident."""
    import torch
    from mse_b200 import synthetic
    d = synthetic.make_dense_corpus(N_DOCS, seed=SEED, device=dev, dtype=torch.bfloat16, chunks_per_doc=chunks_per_doc)
    nat.dense_load(d.emb, d.doc_chunk_off)
    del d
    torch.cuda.empty_cache()
    qv = torch.from_numpy(synthetic.make_query_vectors(BATCH, seed=99)).to(dev)
    cand_off = (torch.arange(BATCH + 1, device=dev, dtype=torch.int32) * TOP_K).contiguous()

    bm_out = (torch.empty((BATCH, TOP_K), dtype=torch.int32, device=dev), torch.empty((BATCH, TOP_K), dtype=torch.float32, device=dev),
              torch.empty((BATCH,), dtype=torch.int32, device=dev))

    def step(i):
        q_off, q_term, q_tf = dev_batches[i % len(dev_batches)]
        doc, score, count = nat.bm25_search(q_off, q_term, q_tf, TOP_K, 0.0, out=bm_out)
        return nat.rerank(cand_off, doc.view(-1), score.view(-1), qv, None, 0.15, 10, max_out), count

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    nat.set_option("reset_timers", 1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        out, count = step(5 + i)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[steps]) / steps
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    ms_median = float(np.median(per_step))
    rr_ms, n = nat.kernel_time("rerank")
    rr_ms /= max(n, 1)
    # stage breakdown (same work, timed separately)
    q_off, q_term, q_tf = dev_batches[0]
    doc, score, count = nat.bm25_search(q_off, q_term, q_tf, TOP_K, 0.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        nat.bm25_search(q_off, q_term, q_tf, TOP_K, 0.0)
    torch.cuda.synchronize()
    bm25_call_ms = (time.perf_counter() - t0) * 1e3 / steps
    t0 = time.perf_counter()
    for i in range(steps):
        nat.rerank(cand_off, doc.view(-1), score.view(-1), qv, None, 0.15, 10, max_out)
    torch.cuda.synchronize()
    rerank_call_ms = (time.perf_counter() - t0) * 1e3 / steps
    rows = float(out[5].float().mean().item())
    cands = float(count.float().mean().item())
    alg = BATCH * (2.0 * 768 * rows + 12.0 * cands + 8.0 * max_out)
    return {"workload": f"hybrid: BM25 top-{TOP_K} over {N_DOCS} docs -> rerank <=10 of {chunks_per_doc} chunks/doc (768-d bf16) -> top-{max_out}, batch {BATCH}",
            "ms_per_batch": ms, "ms_per_batch_median": ms_median, "ms_per_batch_max": float(np.max(per_step)), "slowest_step": int(np.argmax(per_step)), "hybrid_queries_per_s": BATCH / (ms / 1e3), "rerank_kernel_ms": rr_ms,
            "bm25_call_ms": bm25_call_ms, "rerank_call_ms": rerank_call_ms,
            "rerank_rows_per_query": rows, "rerank_GBps_algorithmic": alg / (rr_ms * 1e-3) / 1e9,
            "rerank_frac_of_hbm_peak": alg / (rr_ms * 1e-3) / 1e9 / peak}


def doc_sharded_supplement(args, c, rank, world, dev, local_rank, steps=20, warmup=3):
    """The N x 1024-query batch against the corpus sharded by doc range over the same N GPUs (the layout an index that
    does not fit one GPU needs): shard-local search cut to m entries, NCCL all-to-all to the query-owning rank, merge,
    exact repeat when a cut could hide a result — and a bit-for-bit check against all-gather + merge everywhere."""
    import torch
    import torch.distributed as dist
    from mse_b200 import _native, synthetic
    from mse_b200.sharding import ShardedSearcher
    term_off, post_doc, post_tf, doc_len, doc_base = shard_corpus(c, rank, world)
    nat = _native.NativeIndex(local_rank)
    nat.bm25_load(term_off, post_doc, post_tf, doc_len, c.idf, c.avgdl, doc_base=doc_base)
    searcher = ShardedSearcher(nat, rank, world)
    GB = BATCH * world
    m_local = max(1, min(TOP_K, int(args.exchange_slack * TOP_K / world) + 32))
    batches = [tuple(torch.from_numpy(a).to(dev) for a in synthetic.make_bm25_queries(c, GB, seed=SEED + 7001 + i))
               for i in range(steps + warmup)]
    run = lambda b: searcher.bm25_search_owner(b[0], b[1], b[2], TOP_K, 0.0, slack=args.exchange_slack)
    for i in range(warmup):
        run(batches[i])
    dist.barrier(); torch.cuda.synchronize()
    nat.set_option("reset_timers", 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        o_doc, o_score, o_count = run(batches[warmup + i])
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    score_ms, n = nat.kernel_time("bm25_score")
    select_ms, _ = nat.kernel_time("topk_select")
    prep_ms, _ = nat.kernel_time("bm25_prepare")
    b = batches[warmup]
    o_doc, o_score, o_count = run(b)
    a_doc, a_score, a_count = searcher.bm25_search(b[0], b[1], b[2], TOP_K, 0.0)
    blk = slice(rank * BATCH, (rank + 1) * BATCH)
    valid = torch.arange(TOP_K, device=dev).unsqueeze(0) < o_count.unsqueeze(1)
    same = torch.equal(o_count, a_count[blk]) and torch.equal(o_doc[valid], a_doc[blk][valid]) and \
        torch.equal(o_score[valid], a_score[blk][valid])
    d = torch.tensor([0 if same else 1], device=dev)
    dist.all_reduce(d, op=dist.ReduceOp.SUM)
    nat.close()
    return {"workload": f"{GB}-query batch replicated, corpus in {world} doc-range shards, NCCL all-to-all of {m_local}-entry shard "
                        f"lists to the query-owning rank, exact merge",
            "queries_per_s": GB / (ms / 1e3), "ms_per_step": ms, "steps": steps,
            "score_ms": score_ms / max(1, n), "select_ms": select_ms / max(1, n), "prepare_ms": prep_ms / max(1, n),
            "exchange_fallback_steps": int(getattr(searcher, "fallbacks", 0)),
            "ranks_differing_from_allgather_merge": int(d.item())}


def run_reference_arm(args, rank, world):
    """The reference's CPU path for this workload: the oracle port of BM25.search's Python loop
    (the reference is pure Python; nothing compiles to oracle/_ref), one process per host core."""
    if rank != 0:
        return
    import multiprocessing as mp
    import torch
    from mse_b200 import synthetic
    from oracle import bm25_oracle as bo
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    n_docs = N_DOCS if dev != "cpu" else 50_000          # CPU-only container: tiny corpus just to exercise the arm
    c = synthetic.make_bm25_corpus(n_docs, vocab=VOCAB, seed=SEED, device=dev)
    global _REF_IX, _REF_Q
    _REF_IX = bo.Bm25Arrays(c.term_off.cpu().numpy(), c.post_doc.cpu().numpy(), c.post_tf.cpu().numpy(),
                            c.doc_len.cpu().numpy(), c.idf.cpu().numpy(), c.avgdl, c.total_docs, c.doc_ids.cpu().numpy())
    q_off, q_term, q_tf = synthetic.make_bm25_queries(c, BATCH, seed=SEED + 1)
    _REF_Q = [[int(t) for s in range(q_off[i], q_off[i + 1]) for t in [q_term[s]] * int(q_tf[s])] for i in range(BATCH)]
    cores = os.cpu_count() or 1
    per_step = min(BATCH, 2 * cores)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_ref_one, range(min(cores, BATCH)))                       # warm the workers
        for w in range(args.warmup):
            pool.map(_ref_one, [(w * per_step + j) % BATCH for j in range(per_step)])
        t0 = time.perf_counter()
        for s in range(args.steps):
            pool.map(_ref_one, [((args.warmup + s) * per_step + j) % BATCH for j in range(per_step)])
        dt = time.perf_counter() - t0
    value = args.steps * per_step / dt
    sample = f"{per_step} of the {BATCH} queries per step, full {n_docs}-doc index, faithful Python-loop port"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "queries_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


_REF_IX = None
_REF_Q = None


def _ref_one(i):
    from oracle import bm25_oracle as bo
    return len(bo.search_faithful(_REF_IX, _REF_Q[i], top_k=TOP_K, min_score=0.0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=128, help="queries timed by the CPU baseline (~10-15 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--range-docs", type=int, default=0)
    ap.add_argument("--queries-per-item", type=int, default=0)
    ap.add_argument("--cand-cap", type=int, default=0)
    ap.add_argument("--no-tau", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the supplementary dense-scan (C3) measurements")
    ap.add_argument("--layout", default="auto", choices=["auto", "replicated", "doc-sharded"],
                    help="N > 1: replicate the index and shard the queries, or shard the corpus by doc range (auto: replicate "
                         "when the index fits a quarter of one GPU's memory)")
    ap.add_argument("--exchange", default="owner", choices=["owner", "allgather"],
                    help="N > 1: query-owner merge of an N x 1024 batch (weak scaling) or all-gather + merge of a 1024 batch")
    ap.add_argument("--exchange-slack", type=float, default=2.0, help="owner exchange: shard list length = slack*k/N + 32")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import mse_b200  # noqa: F401
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from mse_b200 import _native, synthetic
    from mse_b200.sharding import ShardedSearcher

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- corpus + index ------------------------------------------------------------------------
    t0 = time.perf_counter()
    c = gen_corpus(dev)
    index_bytes = 8 * int(c.n_postings)
    layout = args.layout
    if layout == "auto":
        layout = "replicated" if index_bytes * 4 <= torch.cuda.get_device_properties(dev).total_memory else "doc-sharded"
    replicated = world > 1 and layout == "replicated"
    full_corpus = c
    if replicated:                                   # every rank is a whole single-GPU engine with its own query batches
        job_world, job_rank, world, rank = world, rank, 1, 0
    else:
        job_world, job_rank = world, rank
    term_off, post_doc, post_tf, doc_len, doc_base = shard_corpus(c, rank, world)
    nat = _native.NativeIndex(local_rank)
    nat.bm25_load(term_off, post_doc, post_tf, doc_len, c.idf, c.avgdl, doc_base=doc_base)
    for name, v in (("bm25_range_docs", args.range_docs), ("bm25_queries_per_item", args.queries_per_item),
                    ("bm25_cand_cap", args.cand_cap)):
        if v:
            nat.set_option(name, v)
    if args.no_tau:
        nat.set_option("bm25_use_tau", 0)
    searcher = ShardedSearcher(nat, rank, world)
    df_global = torch.diff(c.term_off).cpu().numpy()
    n_postings_local = int(post_doc.numel())
    setup_s = time.perf_counter() - t0

    # ---- query batches: a different batch every step (no reuse of a step's postings in L2) -------
    owner = world > 1 and args.exchange == "owner"
    GB = BATCH * world if owner else BATCH                  # queries per step over the whole job (replicated: per rank)
    m_local = max(1, min(TOP_K, int(args.exchange_slack * TOP_K / world) + 32)) if owner else TOP_K
    n_batches = args.steps + args.warmup
    host_batches, dev_batches, postings_per_batch = [], [], []
    for i in range(n_batches):
        q_off, q_term, q_tf = synthetic.make_bm25_queries(c, GB, seed=SEED + 1 + i + (100_000 * job_rank if replicated else 0))
        host_batches.append((q_off, q_term, q_tf))
        dev_batches.append(tuple(torch.from_numpy(a).to(dev) for a in (q_off, q_term, q_tf)))
        postings_per_batch.append(int(df_global[q_term].sum()))
    out = (torch.empty((BATCH, TOP_K), dtype=torch.int32, device=dev), torch.empty((BATCH, TOP_K), dtype=torch.float32, device=dev),
           torch.empty((BATCH,), dtype=torch.int32, device=dev))

    def step_device(i):
        q_off, q_term, q_tf = dev_batches[i]
        if world == 1:
            return nat.bm25_search(q_off, q_term, q_tf, TOP_K, 0.0, out=out)
        if owner:
            return searcher.bm25_search_owner(q_off, q_term, q_tf, TOP_K, 0.0, slack=args.exchange_slack)
        return searcher.bm25_search(q_off, q_term, q_tf, TOP_K, 0.0)

    def sync_all():
        if job_world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if job_world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: inputs resident in HBM ----------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    sync_all()
    nat.set_option("reset_timers", 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active.set()
    e0.record()
    for s in range(args.steps):
        res = step_device(args.warmup + s)
    e1.record()
    sync_all()
    sampler.active.clear()
    ms = max_over_ranks(e0.elapsed_time(e1))
    job_q = GB * (job_world if replicated else 1)           # queries per step over the whole job
    score_ms, score_n = nat.kernel_time("bm25_score")
    select_ms, _ = nat.kernel_time("topk_select")
    prep_ms, _ = nat.kernel_time("bm25_prepare")
    stats = nat.bm25_stats()
    value = args.steps * job_q / (ms / 1000.0)

    # ---- e2e: host buffers through the C ABI (H2D + D2H inside the timed region) ----------------------
    pin = lambda a: torch.from_numpy(a).pin_memory()
    host_pinned = [tuple(pin(a) for a in b) for b in host_batches]
    h_out = (torch.empty((BATCH, TOP_K), dtype=torch.int32).pin_memory(), torch.empty((BATCH, TOP_K), dtype=torch.float32).pin_memory(),
             torch.empty((BATCH,), dtype=torch.int32).pin_memory())

    def step_host(i):
        q_off, q_term, q_tf = host_pinned[i]
        if world == 1:
            nat.bm25_search(q_off, q_term, q_tf, TOP_K, 0.0, out=h_out)      # MSE_HOST path: copies + sync inside
        else:
            d = tuple(t.to(dev, non_blocking=True) for t in (q_off, q_term, q_tf))
            if owner:                                                        # this rank's block of the batch comes back
                r = searcher.bm25_search_owner(d[0], d[1], d[2], TOP_K, 0.0, slack=args.exchange_slack)
            else:
                r = searcher.bm25_search(d[0], d[1], d[2], TOP_K, 0.0)
            for dst, src in zip(h_out, r):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step_host(i)
    sync_all()
    sampler.active.set()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_host(args.warmup + s)
    sync_all()
    e2e_s = time.perf_counter() - t0
    sampler.active.clear()
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = args.steps * job_q / e2e_s
    # whole-job bytes: every rank uploads the (replicated) query CSR; with the owner exchange each rank reads back its
    # own 1024-query block, with the all-gather variant every rank reads back the replicated result
    h2d = job_world * int(np.mean([sum(a.nbytes for a in b) for b in host_batches[args.warmup:]]))
    d2h = (BATCH * TOP_K * 8 + BATCH * 4) * job_world
    sampler.stop()

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    peak, peak_src = measured_peaks()
    timed_post = postings_per_batch[args.warmup:]
    # algorithmic bytes per launch: 12 B per posting of THIS rank's shard + 8 B per emitted result
    frac_local = n_postings_local / max(1, int(c.n_postings))
    alg_bytes = 12.0 * float(np.mean(timed_post)) * frac_local + 8.0 * m_local * GB
    avg_score_ms = score_ms / max(1, score_n)
    achieved = alg_bytes / (avg_score_ms * 1e-3) / 1e9 if avg_score_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "bm25_score_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- N > 1: the owner exchange against the all-gather + merge-everywhere exchange on one batch (outside the
    # timed regions; both are exact, so this rank's block must be identical bit for bit) -----------------------
    exchange_check = None
    if owner:
        q_off, q_term, q_tf = dev_batches[args.warmup]
        o_doc, o_score, o_count = searcher.bm25_search_owner(q_off, q_term, q_tf, TOP_K, 0.0, slack=args.exchange_slack)
        a_doc, a_score, a_count = searcher.bm25_search(q_off, q_term, q_tf, TOP_K, 0.0)
        blk = slice(rank * BATCH, (rank + 1) * BATCH)
        valid = torch.arange(TOP_K, device=dev).unsqueeze(0) < o_count.unsqueeze(1)
        same = torch.equal(o_count, a_count[blk]) and torch.equal(o_doc[valid], a_doc[blk][valid]) and \
            torch.equal(o_score[valid], a_score[blk][valid])
        t = torch.tensor([0 if same else 1], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        exchange_check = {"queries": GB, "ranks_differing_from_allgather_merge": int(t.item())}

    # ---- parity spot-check + CPU baseline (rank 0, N=1) -------------------------------------------------
    cpu_baseline = None
    parity = None
    if job_rank == 0 and job_world == 1 and not args.no_cpu_baseline:
        from oracle import bm25_oracle as bo
        ix = bo.Bm25Arrays(c.term_off.cpu().numpy(), c.post_doc.cpu().numpy(), c.post_tf.cpu().numpy(), c.doc_len.cpu().numpy(),
                           c.idf.cpu().numpy(), c.avgdl, c.total_docs, c.doc_ids.cpu().numpy())
        q_off, q_term, q_tf = host_batches[args.warmup]
        n_s = min(args.cpu_sample, BATCH)
        qs = [[int(t) for s in range(q_off[i], q_off[i + 1]) for t in [q_term[s]] * int(q_tf[s])] for i in range(n_s)]
        t0 = time.perf_counter()
        refs = [bo.search_faithful(ix, q, top_k=TOP_K, min_score=0.0) for q in qs]
        faithful_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        fast = [bo.search_fast(ix, q, top_k=TOP_K, min_score=0.0) for q in qs]
        fast_s = time.perf_counter() - t0
        g_doc, g_score, g_count = nat.bm25_search(q_off, q_term, q_tf, TOP_K, 0.0)
        bad, worst, same_ids, n_ids = 0, 0.0, 0, 0
        for i in range(n_s):
            rd = np.asarray([d for d, _ in refs[i]]); rs = np.asarray([s for _, s in refs[i]])
            n = int(g_count[i])
            if n != len(rd):
                bad += 1
                continue
            scale = bo.abs_contrib_sum(ix, qs[i])[rd] if len(rd) else np.zeros(0)
            tol = 1e-5 * np.maximum(np.maximum(np.abs(rs), scale), 1e-30)
            diff = np.abs(g_score[i, :n] - rs)
            worst = max(worst, float((diff / np.maximum(np.maximum(np.abs(rs), scale), 1e-30)).max(initial=0.0)))
            # north_star rule: scores within tolerance; ids identical except where the scores that decide the
            # order are tied inside that tolerance (a different doc at rank r must carry rank r's score)
            bad += int(np.any(diff > tol))
            same_ids += int(np.sum(g_doc[i, :n] == rd)); n_ids += n
        parity = {"queries_checked": n_s, "queries_failing": bad, "max_rel_err": worst, "tolerance": 1e-5,
                  "rank_positions_with_identical_doc_id": same_ids / max(1, n_ids),
                  "rule": "score at every rank within 1e-5 relative (floor: sum of |term contributions|); ids may differ only at such ties"}
        cpu_baseline = {"value": n_s / faithful_s, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"{n_s} of the {BATCH} queries of one batch, full 1M-doc index, faithful Python-loop port of "
                                  f"bm25_indexer.py:435-485 (no SQL cost)",
                        "vectorised_numpy_value": n_s / fast_s}

    # the CPU-baseline leg leaves millions of Python objects behind: collect now and keep the survivors out of later
    # collections, so that no generation-2 pass lands inside a timed supplement loop (one cost ~28 ms when it did)
    import gc
    gc.collect()
    gc.freeze()

    # ---- supplementary: dense exhaustive scan (BASELINE.json configs[2]) on rank 0 at N=1 ------------------
    dense = None
    hybrid = None
    if job_rank == 0 and job_world == 1 and not args.no_dense:
        try:
            hybrid = hybrid_supplement(nat, dev, dev_batches, peak)
        except Exception as e:  # noqa: BLE001 - the headline number must not depend on the supplements
            hybrid = {"error": repr(e)}
        try:
            dense = dense_scan_supplement(nat, dev, peak)
        except Exception as e:  # noqa: BLE001
            dense = {"error": repr(e)}

    # ---- N > 1, replicated layout: the doc-sharded exchange measured on the same GPUs as a supplement ------------
    doc_sharded = None
    if replicated and not args.no_dense:
        try:
            doc_sharded = doc_sharded_supplement(args, full_corpus, job_rank, job_world, dev, local_rank)
        except Exception as e:  # noqa: BLE001 - the headline number must not depend on the supplements
            doc_sharded = {"error": repr(e)}

    if job_rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": job_world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if (owner or replicated or world == 1) else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_docs": N_DOCS, "vocab": VOCAB, "postings": int(c.n_postings), "batch": BATCH,
                       "global_batch": job_q, "top_k": TOP_K, "layout": layout if job_world > 1 else "single GPU",
                       "sharding": (f"index replicated on {job_world} GPUs ({index_bytes / 1e9:.2f} GB of postings fits one), queries sharded: "
                                    f"{BATCH} per rank and step, no data-path collective" if replicated else
                                    "none" if world == 1 else
                                    f"doc-range x{world}, {GB}-query batch replicated, query-owner merge: NCCL all-to-all of "
                                    f"{m_local}-entry shard lists, exact (full-list repeat when a cut could hide a result)"
                                    if owner else f"doc-range x{world}, all-gather + merge on every rank"),
                       "exchange_fallback_steps": int(getattr(searcher, "fallbacks", 0)),
                       "l2_policy": "inputs larger than L2: 1.5 GB index, a different query batch every step",
                       "postings_per_query": float(np.mean(timed_post)) / GB},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": args.steps * (3 + (1 if world > 1 else 0)) * job_world,   # prepare, score, select (+ merge) per rank
            "roofline": {"bound": "hbm", "kernel": "bm25_score_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": avg_score_ms},
            "cpu_baseline": cpu_baseline,
            "clocks": sampler.summary(),
            "breakdown": {"prepare_ms": prep_ms / max(1, score_n), "score_ms": avg_score_ms, "select_ms": select_ms / max(1, score_n),
                          "candidates_emitted_per_query": stats["emitted"] / GB, "rerun_queries": stats["rerun_queries"],
                          "ranges": stats["ranges"], "score_ctas": stats["ctas"], "setup_s": setup_s},
            "parity": parity,
            "exchange_check": exchange_check,
            "doc_sharded": doc_sharded,
            "hybrid": hybrid,
            "dense_scan": dense,
        }
        print(json.dumps(line))
    if job_world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
