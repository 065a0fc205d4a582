"""Document-range sharding across the GPUs of one box.

Each rank owns a contiguous range of the dense doc index: all postings and all chunks of its
documents, with GLOBAL ``idf`` / ``avgdl`` (per-shard statistics would change scores).  A batch is
scored by every rank against its shard; the only exchange step is an all-gather of the per-rank
top-k lists (``B*k*8`` bytes per rank, NCCL over NVLink when the tensors are CUDA, gloo on CPU
tensors in the unit tests) followed by a device-side merge (``mse_topk_merge``) with the same
ordering rule (score descending, ties to the lower doc id).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def all_gather_topk(doc, score, count, group=None):
    """doc/score [B,k], count [B] (torch tensors, CUDA for NCCL) -> stacked [W,B,k], [W,B,k], [W,B]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    g_doc = torch.empty((world,) + tuple(doc.shape), dtype=doc.dtype, device=doc.device)
    g_score = torch.empty((world,) + tuple(score.shape), dtype=score.dtype, device=score.device)
    g_count = torch.empty((world,) + tuple(count.shape), dtype=count.dtype, device=count.device)
    for out, t in ((g_doc, doc), (g_score, score), (g_count, count)):
        if t.is_cuda:
            dist.all_gather_into_tensor(out, t.contiguous(), group=group)      # one NCCL all-gather over NVLink
        else:
            dist.all_gather([out[w] for w in range(world)], t.contiguous(), group=group)   # gloo (CPU tests)
    return g_doc, g_score, g_count


def exchange_topk(doc, score, count, top_k: int, merge, group=None, slack: float = 2.0, extra: int = 32):
    """Exchange step of a doc-sharded top-k: all-gather + merge, exact, with a truncated first attempt.

    With W shards a rank contributes ~top_k/W entries to the global top-k, so the first attempt gathers only the first
    m = slack*top_k/W + extra entries of every rank (W times less data, a W times shorter merge).  The attempt is exact
    unless some rank was cut (it holds more than m entries) and its m-th score is still >= the merged k-th score (or
    the merged list is short): that rank could own more of the top-k, so the full lists are exchanged instead.  The
    decision is taken from all-gathered values and is therefore the same on every rank.
    ``merge(g_doc, g_score, g_count, top_k)`` is ``NativeIndex.topk_merge`` on CUDA tensors."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    k_local = int(doc.shape[1])
    m = max(1, min(k_local, int(slack * top_k / world) + extra))
    if world == 1 or m >= k_local:
        return merge(*all_gather_topk(doc, score, count, group), top_k)
    cut = torch.clamp(count, max=m)
    g_doc, g_score, g_cut = all_gather_topk(doc[:, :m].contiguous(), score[:, :m].contiguous(), cut, group)
    g_full = torch.empty((world,) + tuple(count.shape), dtype=count.dtype, device=count.device)
    if count.is_cuda:
        dist.all_gather_into_tensor(g_full, count.contiguous(), group=group)
    else:
        dist.all_gather([g_full[w] for w in range(world)], count.contiguous(), group=group)
    out_doc, out_score, out_count = merge(g_doc, g_score, g_cut, top_k)
    out_doc, out_score, out_count = (torch.as_tensor(x) for x in (out_doc, out_score, out_count))
    B = int(count.shape[0])
    rows = torch.arange(B, device=out_score.device)
    full = out_count.to(torch.int64) >= top_k
    kth = out_score[rows, torch.clamp(out_count.to(torch.int64) - 1, min=0)]          # merged k-th score (valid where `full`)
    truncated = g_full.to(out_score.device) > m                                        # [W, B]
    mth = g_score[:, :, m - 1].to(out_score.device)                                    # m-th score of every rank
    unsafe = truncated & (~full.unsqueeze(0) | (mth >= kth.unsqueeze(0)))
    if bool(unsafe.any()):
        return merge(*all_gather_topk(doc, score, count, group), top_k)
    return out_doc, out_score, out_count


def cut_could_hide_a_result(g_score, g_count, out_score, out_count, top_k: int):
    """Exactness rule of the truncated exchange.  ``g_score [W, B, m]`` / ``g_count [W, B]`` are shard lists cut to m
    entries, ``out_score [B, top_k]`` / ``out_count [B]`` their merge.  A cut list (it came back full) whose last score
    is still >= the merged k-th score — or any cut list when the merge is short — may hide an entry of the true top-k;
    otherwise every unsent entry scores strictly below k entries that were sent, and the merge is exact.
    Returns a 0-d bool tensor."""
    import torch
    g_score, g_count, out_score, out_count = (torch.as_tensor(x) for x in (g_score, g_count, out_score, out_count))
    m, dev = int(g_score.shape[2]), out_score.device
    oc = out_count.to(torch.int64)
    short = oc < top_k
    kth = out_score[torch.arange(out_score.shape[0], device=dev), torch.clamp(oc - 1, min=0)]   # valid where not `short`
    cut = g_count.to(dev) >= m
    last = g_score[:, :, m - 1].to(dev)
    return (cut & (short.unsqueeze(0) | (last >= kth.unsqueeze(0)))).any()


def _all_to_all(t, group=None):
    """[W, ...] -> [W, ...]: block w of the input goes to rank w; block w of the output came from rank w."""
    import torch
    import torch.distributed as dist
    out = torch.empty_like(t)
    dist.all_to_all_single(out, t.contiguous(), group=group)          # NCCL over NVLink on CUDA tensors, gloo on CPU
    return out


def exchange_topk_owner(doc, score, count, top_k: int, merge, group=None):
    """Exchange step of a doc-sharded top-k when the batch is too large to merge everywhere: QUERY-OWNER merge.

    ``doc/score [W*Bq, m]``, ``count [W*Bq]`` are this rank's shard-local lists for the whole replicated batch, cut to
    m <= top_k entries per query.  Block w (queries w*Bq .. (w+1)*Bq) is sent to rank w, so rank r receives the W shard
    lists of ITS Bq queries ([W, Bq, m] — the input layout of ``mse_topk_merge``) and merges them to the global top_k:
    a reduce-scatter-shaped version of the all-gather + merge, moving and merging W times less per rank.

    Returns ``(doc [Bq, top_k], score, count, unsafe)``.  ``unsafe`` is a 0-d bool tensor, identical on every rank
    (all-reduced): some shard list was cut (it came back full, m < top_k) while its last score was still >= the merged
    k-th score, or the merged list is short — that shard may own more of the top-k than it sent, and the caller must
    repeat the step with m == top_k, which is always exact."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n, m = int(doc.shape[0]), int(doc.shape[1])
    assert n % world == 0, "the batch must split evenly over the query owners"
    bq = n // world
    g_doc = _all_to_all(doc.view(world, bq, m), group)
    g_score = _all_to_all(score.view(world, bq, m), group)
    g_count = _all_to_all(count.view(world, bq), group)
    out_doc, out_score, out_count = (torch.as_tensor(x) for x in merge(g_doc, g_score, g_count, top_k))
    if m >= top_k:                                 # nothing was cut (the same decision on every rank: no exchange needed)
        return out_doc, out_score, out_count, torch.zeros((), dtype=torch.bool, device=out_score.device)
    unsafe = cut_could_hide_a_result(g_score, g_count, out_score, out_count, top_k)
    flag = unsafe.to(torch.int32).reshape(1)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    return out_doc, out_score, out_count, flag[0] > 0


def merge_topk_host(g_doc: np.ndarray, g_score: np.ndarray, g_count: np.ndarray, top_k: int):
    """Host statement of the merge rule (used by the gloo tests to check the collective plumbing, and
    as the specification of ``mse_topk_merge``): concatenate valid entries, order by (score desc,
    doc asc), keep top_k."""
    W, B, k = g_doc.shape
    out_doc = np.full((B, top_k), -1, dtype=np.int32)
    out_score = np.zeros((B, top_k), dtype=np.float32)
    out_count = np.zeros(B, dtype=np.int32)
    for q in range(B):
        d = np.concatenate([g_doc[w, q, :g_count[w, q]] for w in range(W)])
        s = np.concatenate([g_score[w, q, :g_count[w, q]] for w in range(W)])
        order = np.lexsort((d, -s.astype(np.float64)))[:top_k]
        n = len(order)
        out_doc[q, :n], out_score[q, :n], out_count[q] = d[order], s[order], n
    return out_doc, out_score, out_count


class ShardedSearcher:
    """Wraps a per-rank ``NativeIndex`` loaded with this rank's doc range."""

    def __init__(self, native, rank: int, world: int, group=None):
        self.native, self.rank, self.world, self.group = native, rank, world, group

    def _merge(self, doc, score, count, top_k):
        if self.world == 1:
            return doc, score, count
        return exchange_topk(doc, score, count, top_k, self.native.topk_merge, self.group)

    def bm25_search(self, q_off, q_term, q_tf, top_k: int, min_score: float = 0.0):
        doc, score, count = self.native.bm25_search(q_off, q_term, q_tf, top_k, min_score)
        return self._merge(doc, score, count, top_k)

    def bm25_search_owner(self, q_off, q_term, q_tf, top_k: int, min_score: float = 0.0, slack: float = 2.0, extra: int = 32):
        """Large replicated batch (a multiple of the world size): every rank scores ALL queries against its shard but
        keeps only m = slack*top_k/W + extra entries per query (a shard contributes ~top_k/W to the global top-k), the
        lists travel to the rank that owns the query (``exchange_topk_owner``) and are merged there.  Exact: when a cut
        could have hidden a top-k entry the step is repeated with m == top_k.  Returns the results of this rank's
        block of the batch: queries ``rank*Bq .. (rank+1)*Bq``."""
        if self.world == 1:
            return self.native.bm25_search(q_off, q_term, q_tf, top_k, min_score)
        m = max(1, min(top_k, int(slack * top_k / self.world) + extra))
        doc, score, count = self.native.bm25_search(q_off, q_term, q_tf, m, min_score)
        o_doc, o_score, o_count, unsafe = exchange_topk_owner(doc, score, count, top_k, self.native.topk_merge, self.group)
        if m < top_k and bool(unsafe):
            self.fallbacks = getattr(self, "fallbacks", 0) + 1
            doc, score, count = self.native.bm25_search(q_off, q_term, q_tf, top_k, min_score)
            o_doc, o_score, o_count, _ = exchange_topk_owner(doc, score, count, top_k, self.native.topk_merge, self.group)
        return o_doc, o_score, o_count

    def hybrid_rerank(self, cand_off, cand_doc, cand_bm25, q, n_docs_global: int, url_group=None,
                      smoothing: float = 0.15, max_chunks: int = 10, max_out: int = 1000):
        """Rerank of replicated (already merged) BM25 candidates against a doc-range-sharded chunk table:
        per-rank cosines -> all-reduce(sum) of the (query, candidate, row) arrays -> fuse on every rank."""
        exchange, survivors = self.native.rerank_shard_cos(cand_off, cand_doc, cand_bm25, q, n_docs_global, url_group, max_chunks)
        if self.world > 1:
            import torch.distributed as dist
            for t in exchange:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return self.native.rerank_shard_fuse(exchange, survivors, smoothing, max_out)

    def dense_scan(self, q, top_k: int):
        doc, score, count = self.native.dense_scan(q, top_k)
        return self._merge(doc, score, count, top_k)
