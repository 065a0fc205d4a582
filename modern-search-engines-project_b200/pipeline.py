"""Batched end-to-end search: the GPU-side counterpart of ``search_api.py``'s batch path
(``/root/reference/search_api.py:155-166,204-367``) — SURVEY.md §8f N1.

``queries.txt`` lines ``query_num<TAB>query_text`` in, lines ``query_num<TAB>rank<TAB>url<TAB>score(.3f)``
out (``:214-235,290,348-353``), but every query of the file goes through ONE ``search_batch`` call and
ONE in-process ``rerank_batch`` call instead of a Python loop with an HTTP hop per query.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .bm25_indexer import BM25
from .reranker import DocumentScore, Reranker, WindowScore, hybrid_diversification

TOP_K_RETRIEVAL = 1000     # config.py:13


def preprocess_query(query: str) -> str:
    """``search_api.py:155-166``: lower-case, tuebingen/tubingen -> tübingen, else append it."""
    query = query.strip().lower()
    if "tuebingen" in query or "tubingen" in query or "tübingen" in query:
        query = query.replace("tuebingen", "tübingen").replace("tubingen", "tübingen")
    else:
        query = f"{query} tübingen"
    return query.replace("tuebingen", "tübingen").replace("tubingen", "tübingen").strip().lower()


def read_queries(path: str) -> List[Tuple[str, str]]:
    out = []
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            parts = line.split("\t")
            if len(parts) >= 2:
                out.append((parts[0].strip(), parts[1].strip()))
    return out


class HybridSearch:
    def __init__(self, bm25: BM25, reranker: Reranker, embed: Optional[Callable[[str], np.ndarray]] = None,
                 top_k_retrieval: int = TOP_K_RETRIEVAL):
        self.bm25, self.reranker, self.embed, self.top_k_retrieval = bm25, reranker, embed or reranker.embed, top_k_retrieval
        # BM25 results are positions in bm25.global_doc_ids; the reranker indexes reranker.doc_ids.  The two are the same
        # array when the Reranker is built with doc_ids=bm25.global_doc_ids; when they differ (a
        # bm25_doc_stats id missing from urlsDB, bm25_indexer.py:506) positions are translated through the ids and
        # candidates unknown to the reranker are dropped, as the reference's JOIN on urlsDB drops them
        # (reranker_api.py:36-47).
        self._remap = None
        bm_ids = np.asarray(getattr(bm25, "global_doc_ids", reranker.doc_ids), dtype=np.int64)
        if not np.array_equal(bm_ids, reranker.doc_ids):
            rr = reranker.doc_ids
            pos = np.minimum(np.searchsorted(rr, bm_ids), max(len(rr) - 1, 0))
            self._remap = np.where(rr[pos] == bm_ids, pos, -1).astype(np.int64) if len(rr) else np.full(len(bm_ids), -1, np.int64)

    def search_batch(self, queries: Sequence[str], query_vecs: Optional[np.ndarray] = None, preprocess: bool = True):
        """Returns, per query, the reranked ``[(doc_id, url, score)]`` list (<= 100 entries)."""
        qs = [preprocess_query(q) for q in queries] if preprocess else list(queries)
        q_off, q_term, q_tf = self.bm25.encode_queries(qs)
        doc, score, count = self.bm25.search_batch_terms(q_off, q_term, q_tf, self.top_k_retrieval, 0.0)
        if query_vecs is None:
            query_vecs = np.stack([self.embed(q) for q in qs])
        cand = [doc[i, :count[i]] for i in range(len(qs))]
        sims = [score[i, :count[i]] for i in range(len(qs))]
        if self._remap is not None:
            for i in range(len(qs)):
                m = self._remap[cand[i]]
                cand[i], sims[i] = m[m >= 0].astype(np.int32), sims[i][m >= 0]
        o_doc, o_score, o_orig, o_chunk, o_count, o_rows = self.reranker.rerank_batch(cand, sims, query_vecs)
        results = []
        for i in range(len(qs)):
            n = int(o_count[i])
            ids = self.reranker.doc_ids[o_doc[i, :n]]
            urls = self.reranker.store.urls([int(x) for x in ids]) if hasattr(self.reranker.store, "urls") else {}
            docs = [DocumentScore(doc_id=str(d), title="", url=urls.get(int(d), "") or "", similarity_score=float(s),
                                  original_similarity=float(o),
                                  most_relevant_window=WindowScore(text="", similarity_score=float(s), doc_id=str(d), title="",
                                                                   window_index=int(c)))
                    for d, s, o, c in zip(ids.tolist(), o_score[i, :n], o_orig[i, :n], o_chunk[i, :n])]
            picked = hybrid_diversification(docs, top_k=self.reranker.top_k) if self.reranker.diversification \
                else docs[:self.reranker.top_k]
            results.append([(int(d.doc_id), d.url, d.similarity_score) for d in picked])
        return results

    def batch_search_file(self, queries_path: str, out_path: Optional[str] = None, query_vecs=None) -> List[str]:
        qs = read_queries(queries_path)
        res = self.search_batch([t for _, t in qs], query_vecs)
        lines = [f"{num}\t{rank}\t{url}\t{score:.3f}" for (num, _), r in zip(qs, res)
                 for rank, (_, url, score) in enumerate(r, start=1)]
        if out_path:
            with open(out_path, "w", encoding="utf-8") as f:
                for ln in lines:
                    f.write(ln + "\n")
        return lines
