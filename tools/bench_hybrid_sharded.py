#!/usr/bin/env python
"""Sharded hybrid query on N GPUs: BM25 over doc-range shards -> NCCL all-gather + merge
(global top-1000, replicated) -> per-rank cosines of owned candidates -> NCCL all-reduce(sum) -> fuse.
Checks on a small corpus that the sharded result equals the unsharded one (rank 0 also holds the whole
index), then times the C2-sized corpus (1M docs, 5M chunks).  Launch with torchrun."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import mse_b200  # noqa
from mse_b200 import _native, synthetic
from mse_b200.sharding import ShardedSearcher

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=1_000_000)
ap.add_argument("--check-docs", type=int, default=60_000)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)


def build(n_docs, vocab, B):
    c = synthetic.make_bm25_corpus(n_docs, vocab=vocab, seed=1234, device=dev)
    d = synthetic.make_dense_corpus(n_docs, seed=1234, device=dev, dtype=torch.bfloat16, chunks_per_doc=5)
    lo, hi = n_docs * rank // world, n_docs * (rank + 1) // world
    keep = (c.post_doc >= lo) & (c.post_doc < hi)
    df = torch.diff(c.term_off)
    term_of = torch.repeat_interleave(torch.arange(c.n_terms, device=dev), df)
    toff = torch.zeros(c.n_terms + 1, dtype=torch.int64, device=dev)
    toff[1:] = torch.cumsum(torch.bincount(term_of[keep], minlength=c.n_terms), 0)
    nat = _native.NativeIndex(local)
    nat.bm25_load(toff, (c.post_doc[keep] - lo).contiguous(), c.post_tf[keep].contiguous(), c.doc_len[lo:hi].contiguous(), c.idf, c.avgdl, doc_base=lo)
    off = d.doc_chunk_off
    nat.dense_load(d.emb[int(off[lo]):int(off[hi])].contiguous(), (off[lo:hi + 1] - off[lo]).contiguous(), doc_base=lo, chunk_base=int(off[lo]))
    q_off, q_term, q_tf = synthetic.make_bm25_queries(c, B, seed=4321)
    qb = tuple(torch.from_numpy(x).to(dev) for x in (q_off, q_term, q_tf))
    qv = torch.from_numpy(synthetic.make_query_vectors(B, seed=5)).to(dev)
    return c, d, nat, qb, qv


def hybrid(srch, qb, qv, n_docs, top_k=1000, max_out=100):
    doc, score, count = srch.bm25_search(qb[0], qb[1], qb[2], top_k, 0.0)
    B = doc.shape[0]
    cand_off = (torch.arange(B + 1, device=dev, dtype=torch.int32) * top_k).contiguous()
    return srch.hybrid_rerank(cand_off, doc.reshape(-1).contiguous(), score.reshape(-1).contiguous(), qv, n_docs, None, 0.15, 10, max_out)


# ---- correctness at a small size --------------------------------------------------------------------------
c, d, nat, qb, qv = build(a.check_docs, 20_000, 64)
out = hybrid(ShardedSearcher(nat, rank, world), qb, qv, a.check_docs)
torch.cuda.synchronize()
ok = True
if rank == 0:
    whole = _native.NativeIndex(local)
    whole.bm25_load(c.term_off, c.post_doc, c.post_tf, c.doc_len, c.idf, c.avgdl)
    whole.dense_load(d.emb, d.doc_chunk_off)
    ref = hybrid(ShardedSearcher(whole, 0, 1), qb, qv, a.check_docs)
    torch.cuda.synchronize()
    ok = all(bool(torch.equal(x, y)) for x, y in zip(out, ref))
    whole.close()
del c, d
nat.close()
torch.cuda.empty_cache()
# ---- timing at the C2 size ------------------------------------------------------------------------------------
c, d, nat, qb, qv = build(a.docs, 200_000, a.batch)
srch = ShardedSearcher(nat, rank, world)
for _ in range(3):
    hybrid(srch, qb, qv, a.docs)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    out = hybrid(srch, qb, qv, a.docs)
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t[0])
    print(json.dumps({"workload": f"sharded hybrid: BM25 top-1000 over {a.docs} docs -> rerank (5 chunks/doc) -> top-100, batch {a.batch}",
                      "n_gpus": world, "ms_per_batch": ms, "hybrid_queries_per_s": a.batch / (ms / 1e3),
                      "sharded_equals_unsharded_on_check_corpus": ok}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
