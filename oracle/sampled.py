"""TEST INFRASTRUCTURE — parity checks of the GPU path against the CPU oracle on SAMPLED queries of a full-size corpus.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s parity / CPU-baseline legs may import this.

The full-size corpora (``BASELINE.json`` configs[1..4]: 1M-10M docs, 10M-100M chunks) live on the GPU.  The oracle
(``bm25_oracle`` / ``rerank_oracle``: restatements of ``/root/reference/indexer/bm25_indexer.py:383-485`` and
``/root/reference/reranker/reranker_api.py:27-63,273-372``) needs only what the sampled queries touch: the posting
lists of their terms, the document lengths, and the chunk rows of their candidates.  The helpers here copy exactly
those pieces to the host and apply the north_star comparison rule (scores within tolerance; ids identical except
where the scores that decide the order tie inside it).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import bm25_oracle as bo
from . import rerank_oracle as ro


def query_term_lists(q_off, q_term, q_tf, sample: Sequence[int]) -> List[List[int]]:
    """The oracle's query form (a term id repeated qtf times, in slot order) for the sampled queries of a CSR batch."""
    out = []
    for i in sample:
        out.append([int(t) for s in range(int(q_off[i]), int(q_off[i + 1])) for t in [int(q_term[s])] * int(q_tf[s])])
    return out


def bm25_subindex(term_off, post_doc, post_tf, doc_len, idf, avgdl: float, total_docs: float, terms: Sequence[int]) -> bo.Bm25Arrays:
    """``Bm25Arrays`` over the WHOLE document space holding only the posting lists of ``terms`` (every other list is
    left empty: the oracle never looks at it).  ``term_off`` / ``post_*`` / ``doc_len`` / ``idf`` may be torch CUDA
    tensors; only the needed slices are copied."""
    def host(x):
        return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)
    t_off = host(term_off).astype(np.int64)
    need = sorted(set(int(t) for t in terms if 0 <= int(t) < len(t_off) - 1))
    df = np.zeros(len(t_off) - 1, dtype=np.int64)
    for t in need:
        df[t] = t_off[t + 1] - t_off[t]
    new_off = np.zeros(len(t_off), dtype=np.int64)
    new_off[1:] = np.cumsum(df)
    pd = np.empty(int(new_off[-1]), dtype=np.int32)
    pt = np.empty(int(new_off[-1]), dtype=np.int32)
    for t in need:
        a, e = int(t_off[t]), int(t_off[t + 1])
        pd[new_off[t]:new_off[t + 1]] = host(post_doc[a:e])
        pt[new_off[t]:new_off[t + 1]] = host(post_tf[a:e])
    dl = host(doc_len).astype(np.int32)
    return bo.Bm25Arrays(new_off, pd, pt, dl, host(idf).astype(np.float32), float(avgdl), float(total_docs),
                         np.arange(1, len(dl) + 1, dtype=np.int64))


def compare_topk(got_doc, got_score, ref_doc, ref_score, rtol: float, scale=None, atol: float = 0.0) -> Optional[str]:
    """north_star rule; returns None when the lists agree, else a one-line description of the first violation."""
    got_doc = np.asarray(got_doc); ref_doc = np.asarray(ref_doc)
    got_score = np.asarray(got_score, dtype=np.float64); ref_score = np.asarray(ref_score, dtype=np.float64)
    if len(got_doc) != len(ref_doc):
        return f"length {len(got_doc)} != {len(ref_doc)}"
    if len(ref_doc) == 0:
        return None
    tol = rtol * np.maximum(np.abs(ref_score), 0 if scale is None else np.asarray(scale)) + atol
    bad = np.flatnonzero(np.abs(got_score - ref_score) > tol + 1e-300)
    if len(bad):
        i = int(bad[0])
        return f"score at rank {i}: {got_score[i]!r} vs {ref_score[i]!r} (tol {tol[i]:.3g})"
    mism = np.flatnonzero(got_doc != ref_doc)
    if len(mism):
        # a different doc at rank i must carry rank i's score (checked above); the SETS may differ only at the cut-off score
        diff = set(got_doc.tolist()) ^ set(ref_doc.tolist())
        cut = ref_score[-1]
        g_of = dict(zip(got_doc.tolist(), got_score.tolist())); r_of = dict(zip(ref_doc.tolist(), ref_score.tolist()))
        for d in diff:
            s = g_of.get(d, r_of.get(d))
            if abs(s - cut) > 2 * tol[-1] + atol:
                return f"doc {d} (score {s!r}) is in one list only and not at the cut-off score {cut!r}"
    return None


def check_bm25(ix: bo.Bm25Arrays, queries: List[List[int]], got_doc, got_score, got_count, top_k: int, min_score: float = 0.0,
               rtol: float = 1e-5) -> Dict:
    """GPU BM25 lists of the sampled queries vs ``bm25_oracle.search_fast`` (bit-identical to the faithful loop)."""
    failing, worst, same, total = [], 0.0, 0, 0
    for j, terms in enumerate(queries):
        ref = bo.search_fast(ix, terms, top_k=top_k, min_score=min_score)
        rd = np.asarray([d for d, _ in ref], dtype=np.int64); rs = np.asarray([s for _, s in ref], dtype=np.float64)
        n = int(got_count[j])
        scale = bo.abs_contrib_sum(ix, terms)[rd] if len(rd) else np.zeros(0)
        msg = compare_topk(np.asarray(got_doc[j][:n]), np.asarray(got_score[j][:n]), rd, rs, rtol, scale=scale)
        if msg:
            failing.append((j, msg))
        elif len(rd):
            den = np.maximum(np.maximum(np.abs(rs), scale), 1e-30)
            worst = max(worst, float((np.abs(np.asarray(got_score[j][:n], dtype=np.float64) - rs) / den).max()))
            same += int(np.sum(np.asarray(got_doc[j][:n]) == rd)); total += n
    return {"queries_checked": len(queries), "queries_failing": len(failing), "first_failures": [f"q{j}: {m}" for j, m in failing[:3]],
            "max_rel_err": worst, "tolerance": rtol, "rank_positions_with_identical_doc_id": same / max(1, total),
            "rule": "score at every rank within rtol (floor: sum of |term contributions|); ids differ only at such ties"}


def dense_subtable(fetch_rows: Callable[[np.ndarray], np.ndarray], doc_chunk_off, docs: np.ndarray, max_rows_per_doc: int = 10
                   ) -> Tuple[ro.DenseArrays, np.ndarray]:
    """Compact ``DenseArrays`` over the (ascending, unique) global docs ``docs``: rows fetched through
    ``fetch_rows(global_row_ids) -> float32 [n, 768]`` (a D2H gather), chunk ids = global row numbers.  Only the first
    ``max_rows_per_doc`` rows of a doc are fetched (the rerank never reads more, reranker_api.py:58)."""
    def host(x):
        return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)
    docs = np.asarray(docs, dtype=np.int64)
    off = doc_chunk_off
    a = host(off[docs]).astype(np.int64) if hasattr(off, "detach") else np.asarray(off)[docs].astype(np.int64)
    e = host(off[docs + 1]).astype(np.int64) if hasattr(off, "detach") else np.asarray(off)[docs + 1].astype(np.int64)
    n = np.minimum(e - a, max_rows_per_doc)
    new_off = np.zeros(len(docs) + 1, dtype=np.int64)
    new_off[1:] = np.cumsum(n)
    rows = np.concatenate([np.arange(x, x + k, dtype=np.int64) for x, k in zip(a, n)]) if len(docs) else np.zeros(0, np.int64)
    emb = fetch_rows(rows) if len(rows) else np.zeros((0, 768), np.float32)
    return ro.DenseArrays(np.asarray(emb, dtype=np.float32), rows, new_off, docs), docs


def oracle_rerank_global(fetch_rows, doc_chunk_off, cand_doc, cand_score, q, smoothing: float = 0.15, max_chunks: int = 10,
                         faithful: bool = False, url_group=None):
    """``rerank_oracle.rerank`` on a full-size corpus: candidates are GLOBAL doc indices; returns (doc, score, orig,
    chunk, rows) with global doc indices, or None (the reference's HTTP 401)."""
    cand_doc = np.asarray(cand_doc, dtype=np.int64); cand_score = np.asarray(cand_score, dtype=np.float64)
    if url_group is not None and len(cand_doc):                  # URL-group dedupe: lowest doc of a group survives (:38-47)
        grp = np.asarray(url_group)[cand_doc]
        order = np.lexsort((cand_doc, grp))
        keep = np.ones(len(cand_doc), dtype=bool)
        keep[order[1:]] = grp[order[1:]] != grp[order[:-1]]
        cand_doc, cand_score = cand_doc[keep], cand_score[keep]
    docs = np.unique(cand_doc)
    dense, docs = dense_subtable(fetch_rows, doc_chunk_off, docs, max_chunks)
    local = np.searchsorted(docs, cand_doc)
    res = ro.rerank(dense, local, cand_score, np.asarray(q, dtype=np.float32), smoothing, max_chunks, faithful=faithful)
    if res is None:
        return None
    return docs[res.doc], res.score, res.orig, res.best_chunk, res.total_rows


def oracle_hybrid(ix: bo.Bm25Arrays, fetch_rows, doc_chunk_off, terms: List[int], q, top_k: int = 1000, max_out: int = 100,
                  faithful: bool = False, url_group=None):
    """The reference's per-query hybrid path (search_api.py:252-283): BM25 top_k -> rerank -> first max_out."""
    ref = (bo.search_faithful if faithful else bo.search_fast)(ix, terms, top_k=top_k, min_score=0.0)
    if not ref:
        return np.zeros(0, np.int64), np.zeros(0), 0
    out = oracle_rerank_global(fetch_rows, doc_chunk_off, [d for d, _ in ref], [s for _, s in ref], q, faithful=faithful,
                               url_group=url_group)
    if out is None:
        return np.zeros(0, np.int64), np.zeros(0), 0
    return out[0][:max_out], out[1][:max_out], out[4]


def check_hybrid(ix: bo.Bm25Arrays, fetch_rows, doc_chunk_off, queries: List[List[int]], q_vecs, got_doc, got_score, got_count,
                 top_k: int = 1000, max_out: int = 100, atol: float = 3e-3, url_group=None) -> Dict:
    """GPU hybrid results of the sampled queries vs the oracle pipeline.  Fused scores live in [0, 1]; tolerance is
    absolute (the dense term carries the 2e-3 relative bf16/fp32 budget of north_star through the min-max)."""
    failing, worst, same, total = [], 0.0, 0, 0
    for j, terms in enumerate(queries):
        rd, rs, _ = oracle_hybrid(ix, fetch_rows, doc_chunk_off, terms, q_vecs[j], top_k, max_out, url_group=url_group)
        n = int(got_count[j])
        msg = compare_topk(np.asarray(got_doc[j][:n]), np.asarray(got_score[j][:n]), rd, rs, 0.0, atol=atol)
        if msg:
            failing.append((j, msg))
        elif len(rd):
            worst = max(worst, float(np.abs(np.asarray(got_score[j][:n], dtype=np.float64) - rs).max()))
            same += int(np.sum(np.asarray(got_doc[j][:n]) == rd)); total += n
    return {"queries_checked": len(queries), "queries_failing": len(failing), "first_failures": [f"q{j}: {m}" for j, m in failing[:3]],
            "max_abs_err": worst, "tolerance_abs": atol, "rank_positions_with_identical_doc_id": same / max(1, total),
            "rule": "fused score at every rank within atol; ids differ only where the deciding scores tie inside it"}


def dense_scan_slabwise(fetch_slab: Callable[[int, int], np.ndarray], n_chunks: int, doc_chunk_off: np.ndarray, qs: np.ndarray,
                        top_k: int, slab: int = 1 << 20):
    """``rerank_oracle.dense_scan`` for several queries over a table too large to convert at once: float32 slabs
    (``fetch_slab(lo, hi) -> float32 [hi-lo, 768]``, the STORED values), scores accumulated in float64 per slab as the
    oracle does, per-doc max, top-k with ties to the lower doc.  Returns [(doc int64[k], score float64[k])] per query."""
    qs64 = np.asarray(qs, dtype=np.float64)
    off = np.asarray(doc_chunk_off, dtype=np.int64)
    scores = np.empty((len(qs64), n_chunks), dtype=np.float64)
    for a in range(0, n_chunks, slab):
        e = min(n_chunks, a + slab)
        scores[:, a:e] = (fetch_slab(a, e).astype(np.float64) @ qs64.T).T
    has = np.flatnonzero(off[1:] > off[:-1])
    out = []
    for j in range(len(qs64)):
        best = np.maximum.reduceat(scores[j], off[has]) if has.size else np.zeros(0)
        order = np.argsort(-best, kind="stable")[:top_k]
        out.append((has[order].astype(np.int64), best[order]))
    return out
