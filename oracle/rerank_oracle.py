"""TEST INFRASTRUCTURE — CPU restatement of the reference dense rerank path and of the
(removed) exhaustive dense retriever.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.

Follows ``/root/reference/reranker/reranker_api.py`` (config ``reranker/config.yaml:25-30``:
batch_size 32, smoothing 0.15, diversification True, top_k 100):
  * candidate fetch      ``:27-63``    URL-dedupe (MIN(id) per url-up-to-'?'), first <=10 chunks/doc
  * cosine               ``:273-287``  sklearn ``cosine_similarity`` in 32-row batches, float32
  * min-max              ``:289-296``  over all candidate rows; all-equal -> 0.0
  * fusion               ``:360-362``  ``new*(1-smoothing) + old*smoothing``
  * positional weighting ``:299-334``  best chunk += 0.1 - (0.1+0.05)*pos/(n-1), clamp [0,1]
  * per-doc max + sort   ``:370-372``
  * diversification      ``:170-236``
The exhaustive scan (``dense_scan``) reconstructs ``Retriever.quick_search`` whose source is
absent from the reference snapshot (call-site remnants: ``search_api.py:60,87``,
``indexer/embedder.py:54-61``, ``indexer/indexer.py:165``): inner product of the L2-normalised
query with L2-normalised chunk embeddings, max-pooled per document, top-k unique docs, ties to
the lower doc id.

Parity status: rerank PINNED against the unmodified reference ``rerank()`` run under
``oracle/stub_harness.py`` (fixtures in ``tests/golden/``, script ``oracle/make_golden.py``);
``dense_scan`` is a reconstruction — "parity unpinned" for that one function (no reference code
exists to pin it against).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence
from urllib.parse import urlparse

import numpy as np

SMOOTHING = 0.15        # reranker/config.yaml:28
MAX_CHUNKS = 10         # reranker_api.py:58
COS_BATCH = 32          # reranker/config.yaml:27
TOP_K = 100             # reranker/config.yaml:30
MAX_BOOST = 0.1         # reranker_api.py:317
MAX_DECAY = 0.05        # reranker_api.py:318


@dataclass
class DenseArrays:
    emb: np.ndarray                 # float32 [n_chunks, dim]; rows doc-contiguous, ascending chunk id
    chunk_ids: np.ndarray           # int64 [n_chunks]
    doc_chunk_off: np.ndarray       # int64 [N+1] over the dense doc index
    doc_ids: np.ndarray             # int64 [N] ascending
    urls: Optional[List[str]] = None


@dataclass
class RerankRows:
    doc: np.ndarray                 # dense doc index per fetched row
    row: np.ndarray                 # row index into emb
    old: np.ndarray                 # BM25 score per row (float64)


@dataclass
class RerankResult:
    doc: np.ndarray                 # dense doc index, sorted by score desc
    score: np.ndarray               # float64 fused + positional score
    orig: np.ndarray                # float64 min-max'd BM25 score
    best_chunk: np.ndarray          # int64 chunk_id of the representative row
    total_rows: int


def url_key(url: str) -> str:
    """``:44-47`` — url up to (not including) the first '?'."""
    i = url.find("?")
    return url[:i] if i >= 0 else url


def fetch_rows(dense: DenseArrays, cand_doc: Sequence[int], cand_score: Sequence[float],
               max_chunks: int = MAX_CHUNKS) -> RerankRows:
    """``:36-59`` + merge ``:357``.  Representative of a URL group = lowest doc id among the
    candidates; its BM25 score is the representative's own (the merge is on the representative id).
    Frame order: ascending doc, ascending chunk id."""
    cand_doc = np.asarray(cand_doc, dtype=np.int64)
    cand_score = np.asarray(cand_score, dtype=np.float64)
    score_of = {}
    for d, s in zip(cand_doc.tolist(), cand_score.tolist()):
        score_of.setdefault(d, s)
    docs = sorted(score_of)
    if dense.urls is not None:
        rep = {}
        for d in docs:                     # ascending -> first seen is MIN(id)
            rep.setdefault(url_key(dense.urls[d]), d)
        docs = sorted(rep.values())
    r_doc, r_row, r_old = [], [], []
    for d in docs:
        a, e = int(dense.doc_chunk_off[d]), int(dense.doc_chunk_off[d + 1])
        e = min(e, a + max_chunks)
        for r in range(a, e):
            r_doc.append(d); r_row.append(r); r_old.append(score_of[d])
    return RerankRows(np.asarray(r_doc, dtype=np.int64), np.asarray(r_row, dtype=np.int64),
                      np.asarray(r_old, dtype=np.float64))


def cosine_rows(emb_rows: np.ndarray, q: np.ndarray, faithful: bool = True, batch: int = COS_BATCH) -> np.ndarray:
    """``:273-287`` — float32 cosine.  ``faithful`` calls scikit-learn (the reference's dependency)
    in 32-row batches; otherwise one numpy expression in float32."""
    q = np.asarray(q, dtype=np.float32)
    if faithful:
        from sklearn.metrics.pairwise import cosine_similarity
        out = []
        for i in range(0, len(emb_rows), batch):
            out.extend(cosine_similarity(q.reshape(1, -1), np.asarray(emb_rows[i:i + batch], dtype=np.float32))[0])
        return np.asarray(out, dtype=np.float32)
    e = np.asarray(emb_rows, dtype=np.float32)
    en = np.sqrt(np.einsum("ij,ij->i", e, e))
    qn = np.sqrt(np.dot(q, q))
    return ((e @ q) / (en * qn)).astype(np.float32)


def minmax(x: np.ndarray) -> np.ndarray:
    """``:289-296`` in float64."""
    x = np.asarray(x, dtype=np.float64)
    lo, hi = x.min(), x.max()
    if hi == lo:
        return np.zeros_like(x)
    return (x - lo) / (hi - lo)


def rerank(dense: DenseArrays, cand_doc: Sequence[int], cand_score: Sequence[float], q: np.ndarray,
           smoothing: float = SMOOTHING, max_chunks: int = MAX_CHUNKS, faithful: bool = True
           ) -> Optional[RerankResult]:
    """Steps 1-10 of SURVEY.md Appendix A.2.  Returns None when no rows are fetched (the
    reference answers HTTP 401, ``:348-349``)."""
    rows = fetch_rows(dense, cand_doc, cand_score, max_chunks)
    n = len(rows.doc)
    if n == 0:
        return None
    cos = cosine_rows(dense.emb[rows.row], q, faithful)
    new = minmax(cos.astype(np.float64))
    old = minmax(rows.old)
    new = new * (1 - smoothing) + old * smoothing
    # groups are contiguous because rows are ordered by doc
    starts = np.flatnonzero(np.r_[True, rows.doc[1:] != rows.doc[:-1]])
    ends = np.r_[starts[1:], n]
    for a, e in zip(starts, ends):                 # ``:299-334``
        m = e - a
        if m == 1:
            continue
        best = a + int(np.argmax(new[a:e]))        # first occurrence of the max
        pos = best - a                             # rows of a doc are in ascending chunk id
        ratio = pos / max(1, m - 1)
        adj = MAX_BOOST - (MAX_BOOST + MAX_DECAY) * ratio
        new[best] = max(0.0, min(1.0, new[best] + adj))
    d_doc, d_score, d_orig, d_chunk = [], [], [], []
    for a, e in zip(starts, ends):                 # ``:370-371``
        best = a + int(np.argmax(new[a:e]))
        d_doc.append(int(rows.doc[a])); d_score.append(float(new[best]))
        d_orig.append(float(old[best])); d_chunk.append(int(dense.chunk_ids[rows.row[best]]))
    d_doc = np.asarray(d_doc, dtype=np.int64)
    d_score = np.asarray(d_score, dtype=np.float64)
    order = np.argsort(-d_score, kind="stable")    # ``:372`` (tie order unspecified there; here: lower doc first)
    return RerankResult(d_doc[order], d_score[order], np.asarray(d_orig)[order],
                        np.asarray(d_chunk, dtype=np.int64)[order], n)


def domain_of(url: str) -> str:
    """``:170-176``."""
    try:
        return urlparse(url).netloc.lower()
    except Exception:
        return "defaultdomain"


def diversify(urls: Sequence[str], scores: Sequence[float], relevance_threshold: float = 0.8, top_k: int = TOP_K):
    """``:196-236`` on parallel lists sorted by score descending.  Returns (indices into the input,
    final scores) — the scores of back-filled docs are shifted as at ``:229-233``."""
    n = len(urls)
    dom = [domain_of(u) for u in urls]
    sc = [float(s) for s in scores]
    high_dom = {dom[i] for i in range(n) if sc[i] >= relevance_threshold}
    med_dom = {dom[i] for i in range(n) if sc[i] < relevance_threshold} - high_dom
    high = [i for i in range(n) if sc[i] >= relevance_threshold or dom[i] in high_dom]
    med = [i for i in range(n) if sc[i] < relevance_threshold and dom[i] in med_dom]
    high.sort(key=lambda i: sc[i], reverse=True)
    med.sort(key=lambda i: sc[i], reverse=True)

    def cap(idx):
        seen, keep, drop = set(), [], []
        for i in idx:
            if dom[i] in seen:
                drop.append(i)
            else:
                seen.add(dom[i]); keep.append(i)
        return keep, drop

    kh, dh = cap(high)
    km, dm = cap(med)
    # remaining_slots may be <= 0; the reference slices with it unchanged (``:215-219``)
    final = sorted(kh + km[:top_k - len(kh)], key=lambda i: sc[i], reverse=True)
    rest = sorted(dh + dm, key=lambda i: sc[i], reverse=True)
    out_sc = dict((i, sc[i]) for i in final)
    if len(final) < top_k:
        add = rest[:top_k - len(final)]
        if add:
            delta = sc[add[0]] - sc[final[-1]] + 1e-4
            for i in add:
                out_sc[i] = max(0.0, sc[i] - delta)
            final = final + add
    final = sorted(final, key=lambda i: out_sc[i], reverse=True)
    return final, [out_sc[i] for i in final]


def dense_scan(dense: DenseArrays, q: np.ndarray, top_k: int = 1000, normalize_query: bool = True):
    """Reconstructed ``Retriever.quick_search(..., return_unique_docs=True)``: float64 accumulate of
    the stored embeddings against the (normalised) query, per-doc max, top-k, ties -> lower doc."""
    q = np.asarray(q, dtype=np.float64)
    if normalize_query:
        q = q / np.sqrt(np.dot(q, q))
    s = dense.emb.astype(np.float64) @ q
    off = dense.doc_chunk_off
    has = np.flatnonzero(off[1:] > off[:-1])
    if has.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float64)
    best = np.maximum.reduceat(s, off[has])
    order = np.argsort(-best, kind="stable")[:top_k]
    return has[order].astype(np.int64), best[order]
