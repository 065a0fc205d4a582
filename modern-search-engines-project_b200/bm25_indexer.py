"""Drop-in for ``indexer.bm25_indexer.BM25`` (``/root/reference/indexer/bm25_indexer.py:56-568``)
whose query path runs on the GPU.

Same constructor and method names, argument meaning, return shapes and empty-result behaviour as
the reference class, so ``search_api.py:51,88,252`` work unchanged:

    bm_25 = BM25(cfg.DB_PATH, read_only=True)
    results = bm_25.search(query, top_k=1000)      # [{'doc_id', 'score', 'text_snippet'}, ...]

What differs is where the work happens: the four ``bm25_*`` tables are read ONCE at construction
into a CSR inverted index resident in HBM, and ``search`` / ``search_batch`` run the batched
scoring + top-k kernels behind ``mse_bm25_search_batch``.  Host code keeps only what is string
work in the reference: tokenisation (``:149-155``), the term dictionary lookup (``:412-432``) and
the snippet fetch (``:491-514``).
"""
from __future__ import annotations

import math
from collections import defaultdict
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .store import ArrayStore, Bm25Tables, SqlStore, open_store


def spacy_tokenizer():
    """The reference's tokenizer (``:72-80,149-155``): spaCy ``en_core_web_sm`` lemmas that are alpha,
    non-stop, non-punct.  Raises ImportError where spaCy is not installed."""
    import spacy  # type: ignore
    nlp = spacy.load("en_core_web_sm")

    def tok(text: str) -> List[str]:
        return [t.lemma_.lower() for t in nlp(text) if not t.is_stop and not t.is_punct and t.is_alpha]
    return tok


def whitespace_tokenizer(text: str) -> List[str]:
    return text.split()


def shard_bounds(term_weight_per_doc: np.ndarray, world: int) -> List[int]:
    """Contiguous doc-index ranges balanced by per-doc weight (postings or chunks), SURVEY.md §8e."""
    n = len(term_weight_per_doc)
    cum = np.concatenate([[0], np.cumsum(term_weight_per_doc, dtype=np.int64)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(cum, total * r // world)))
    bounds.append(n)
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def slice_bm25_tables(t: Bm25Tables, lo: int, hi: int) -> Bm25Tables:
    """Postings of docs [lo, hi) re-based to local doc indices; idf / avgdl stay global."""
    keep = (t.post_doc >= lo) & (t.post_doc < hi)
    term_of = np.repeat(np.arange(len(t.term_off) - 1, dtype=np.int64), np.diff(t.term_off))
    off = np.zeros(len(t.term_off), dtype=np.int64)
    np.add.at(off, term_of[keep] + 1, 1)
    return Bm25Tables(t.terms, np.cumsum(off), (t.post_doc[keep] - lo).astype(np.int32), t.post_tf[keep],
                      t.doc_ids[lo:hi], t.doc_len[lo:hi], t.idf, t.total_freq, t.avgdl, t.total_docs)


MAX_QUERY_TERMS = 32          # distinct valid terms per query the score kernel stages (csrc/bm25.cuh kMetaSlots)


class BM25:
    def __init__(self, db_path: Optional[str], k1: float = 1.2, b: float = 0.75, read_only: bool = True, *,
                 store=None, tokenizer: Optional[Callable[[str], List[str]]] = None, device: int = 0,
                 doc_range: Optional[Tuple[int, int]] = None, load: bool = True, cache_path: Optional[str] = None,
                 appended_term: Optional[str] = "tübingen"):
        self.db_path = db_path
        self.k1 = k1
        self.b = b
        self.read_only = read_only
        self.store = store if store is not None else open_store(db_path, read_only=read_only)
        self.conn = getattr(self.store, "conn", None)          # the reference exposes .conn (:69)
        self._tokenizer = tokenizer
        self.device = device
        self.doc_range = doc_range
        self.cache_path = cache_path                      # optional .npz of the CSR arrays (skips the SQL scan)
        self.appended_term = appended_term                # the term search_api.py:160-165 appends to every query
        self.tables: Optional[Bm25Tables] = None
        self.native: Optional[_native.NativeIndex] = None
        self._term_index: Dict[str, int] = {}
        self.doc_base = 0
        if load and self.store.has_table("bm25_term_freq"):
            self.reload()

    # ------------------------------------------------------------------ loading
    def reload(self):
        """(Re)reads the bm25_* tables and uploads the CSR index to HBM."""
        all_ids = self.store.all_doc_ids() if hasattr(self.store, "all_doc_ids") else None
        self.fingerprint = self.store.bm25_fingerprint() if hasattr(self.store, "bm25_fingerprint") else None
        full = None
        if self.cache_path:
            from .store import load_bm25_cache, save_bm25_cache
            # staleness is decided by the fingerprint; the doc count of the cached index is union(urlsDB, doc_stats)
            # and need not equal len(urlsDB)
            full = load_bm25_cache(self.cache_path, fingerprint=self.fingerprint)
        if full is None:
            full = self.store.load_bm25(all_ids)
            if self.cache_path:
                save_bm25_cache(self.cache_path, full, self.fingerprint or "")
        self.global_doc_ids = full.doc_ids
        self._df_global = np.diff(full.term_off)
        t = full
        if self.doc_range is not None:
            lo, hi = self.doc_range
            t = slice_bm25_tables(full, lo, hi)
            self.doc_base = lo
        self.tables = t
        self._term_index = {s: i for i, s in enumerate(t.terms)} if t.terms is not None else {}
        if self.native is None:
            self.native = _native.NativeIndex(self.device)
        self.native.bm25_load(np.ascontiguousarray(t.term_off, dtype=np.int64),
                              np.ascontiguousarray(t.post_doc, dtype=np.int32),
                              np.ascontiguousarray(t.post_tf, dtype=np.int32),
                              np.ascontiguousarray(t.doc_len, dtype=np.int32),
                              np.ascontiguousarray(t.idf, dtype=np.float32),
                              t.avgdl, self.k1, self.b, doc_base=self.doc_base)
        # The term the caller appends to every query is in nearly every document (negative idf): the library keeps its
        # per-document impact class in every posting, so that documents it sinks below the bound are never looked up.  Its
        # default is the negative-idf term with the most postings; name the right one when the dictionary holds it.
        # (Speed only: results never depend on it.)
        j = self._term_index.get(self.appended_term, -1) if self.appended_term else -1
        if j >= 0 and float(t.idf[j]) < 0 and hasattr(self.native, "set_option"):
            try:
                self.native.set_option("bm25_class_term", j)
            except _native.NativeError:
                pass                                      # no dense impact row for it (row budget): the default stays

    def refresh(self) -> bool:
        """Reloads the index when the bm25_* tables have changed since it was read (new ``processed_at`` /
        ``last_updated`` stamps, row counts or corpus statistics — `SqlStore.bm25_fingerprint`); returns whether it
        did.  The reference re-reads the tables on every query, so an index rebuilt by ``index_all.py`` is picked up
        at once there; here a serving process calls this between batches (one cheap aggregate query per table)."""
        if not hasattr(self.store, "bm25_fingerprint") or not self.store.has_table("bm25_term_freq"):
            return False
        if self.tables is not None and self.store.bm25_fingerprint() == getattr(self, "fingerprint", None):
            return False
        self.reload()
        return True

    # ------------------------------------------------------------------ tokenisation (host)
    def _tokenize(self, text: str) -> List[str]:
        if self._tokenizer is None:
            self._tokenizer = spacy_tokenizer()
        return self._tokenizer(text)

    def _query_slots(self, terms: Iterable) -> Tuple[List[int], List[int]]:
        """``:405-432``: unique terms in first-occurrence order + query tf; unknown terms dropped."""
        qtf: Dict[int, int] = {}
        order: List[int] = []
        n_terms = len(self._df_global)
        for t in terms:
            j = self._term_index.get(t, -1) if isinstance(t, str) else int(t)
            if j < 0 or j >= n_terms or self._df_global[j] == 0:
                continue
            if j not in qtf:
                qtf[j] = 0
                order.append(j)
            qtf[j] += 1
        return order, [qtf[j] for j in order]

    def encode_queries(self, queries: Sequence) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """List of query strings (tokenised here) or of term lists -> CSR batch."""
        q_off, q_term, q_tf = [0], [], []
        for q in queries:
            terms = self._tokenize(q) if isinstance(q, str) else q
            o, f = self._query_slots(terms)
            q_term.extend(o); q_tf.extend(f); q_off.append(len(q_term))
        return (np.asarray(q_off, dtype=np.int32), np.asarray(q_term, dtype=np.int32), np.asarray(q_tf, dtype=np.int32))

    # ------------------------------------------------------------------ search
    def search_batch_terms(self, q_off, q_term, q_tf, top_k: int = 1000, min_score: float = 0.0):
        """CSR term-id batch -> (doc index [B,k] int32 (global dense index, -1 padded), score [B,k] fp32,
        count [B]).  Accepts numpy (host) or torch CUDA tensors (device-resident, no copies)."""
        self._require_loaded()
        if not _native._is_torch(q_off) and len(q_off) > 1 and int(np.max(np.diff(q_off))) > MAX_QUERY_TERMS:
            return self._search_long_queries(np.asarray(q_off), np.asarray(q_term), np.asarray(q_tf), top_k, min_score)
        return self.native.bm25_search(q_off, q_term, q_tf, top_k, min_score)

    def _search_long_queries(self, q_off, q_term, q_tf, top_k: int, min_score: float):
        """The kernel takes <= 32 distinct terms per query; the reference's ``search`` accepts any length (a pasted
        paragraph).  A longer query is cut into pieces of <= 32 terms, every piece is scored WITHOUT a cut
        (top_k = all docs it touches, min_score = -inf), the piece scores of a document are summed in piece order on
        the host and the reference's filter / order / slice (:480-485) is applied to the sums.  Short queries of the
        same batch still go through one batched call."""
        B = len(q_off) - 1
        n_terms = np.diff(q_off)
        long_q = np.flatnonzero(n_terms > MAX_QUERY_TERMS)
        short_q = np.flatnonzero(n_terms <= MAX_QUERY_TERMS)
        doc = np.full((B, top_k), -1, np.int32); score = np.zeros((B, top_k), np.float32); count = np.zeros(B, np.int32)
        if len(short_q):
            so = np.zeros(len(short_q) + 1, np.int32); st, sf = [], []
            for j, q in enumerate(short_q):
                st.append(q_term[q_off[q]:q_off[q + 1]]); sf.append(q_tf[q_off[q]:q_off[q + 1]]); so[j + 1] = so[j] + n_terms[q]
            d, s, c = self.native.bm25_search(so, np.concatenate(st).astype(np.int32) if so[-1] else np.zeros(0, np.int32),
                                              np.concatenate(sf).astype(np.int32) if so[-1] else np.zeros(0, np.int32), top_k, min_score)
            doc[short_q], score[short_q], count[short_q] = d, s, c
        n_docs = len(self.tables.doc_len)
        k_all = min(_native.MAX_TOPK, max(1, n_docs))
        for q in long_q:
            lo, hi = int(q_off[q]), int(q_off[q + 1])
            total: Dict[int, float] = {}
            for a in range(lo, hi, MAX_QUERY_TERMS):
                e = min(hi, a + MAX_QUERY_TERMS)
                po = np.asarray([0, e - a], np.int32)
                d, s, c = self.native.bm25_search(po, np.ascontiguousarray(q_term[a:e], np.int32),
                                                  np.ascontiguousarray(q_tf[a:e], np.int32), k_all, -3.0e38)
                if int(c[0]) >= k_all and n_docs > k_all:
                    # the piece may touch more documents than one call can return: fail THIS query only (count 0),
                    # never the batch (search_api's batch path would lose every query of the file)
                    import warnings
                    warnings.warn(f"query {int(q)} has {hi - lo} distinct terms and a {e - a}-term piece touches more than "
                                  f"{k_all} documents: not supported, returning no results for it")
                    total = None
                    break
                for dd, ss in zip(d[0, :int(c[0])].tolist(), s[0, :int(c[0])].tolist()):
                    total[dd] = total.get(dd, 0.0) + ss
            if total is None:
                continue
            keep = sorted(((sc, dd) for dd, sc in total.items() if sc >= min_score), key=lambda x: (-x[0], x[1]))[:top_k]
            count[q] = len(keep)
            for r, (sc, dd) in enumerate(keep):
                doc[q, r], score[q, r] = dd, sc
        return doc, score, count

    def search_batch(self, queries: Sequence, top_k: int = 1000, min_score: float = 0.0):
        """Batched ``search``: returns (doc_ids [B,k] int64 urlsDB ids, scores [B,k], counts [B])."""
        q_off, q_term, q_tf = self.encode_queries(queries)
        doc, score, count = self.search_batch_terms(q_off, q_term, q_tf, top_k, min_score)
        ids = np.where(doc >= 0, self.global_doc_ids[np.maximum(doc, 0)], -1)
        return ids, score, count

    def search(self, query: str, top_k: int = 1000, min_score: float = 0.0) -> List[dict]:
        """Same contract as the reference ``search`` (``:383-514``): ``[]`` when the query has no
        tokens, no known terms or no surviving document; otherwise dicts in rank order."""
        query_terms = self._tokenize(query)
        if not query_terms:
            return []
        slots, _ = self._query_slots(query_terms)
        if not slots:
            return []
        ids, score, count = self.search_batch([query_terms], top_k=top_k, min_score=min_score)
        n = int(count[0])
        if n == 0:
            return []
        top_ids = [int(x) for x in ids[0, :n]]
        details = self.store.documents(top_ids)
        out = []
        for d, s in zip(top_ids, score[0, :n].tolist()):
            if d in details:                                    # ids missing from urlsDB are dropped (:506)
                title, text = details[d]
                text = text or ""
                snip = f"{title or 'N/A'}: {text[:200]}"
                if len(text) > 200:
                    snip += "..."
                out.append({"doc_id": d, "score": s, "text_snippet": snip})
        return out

    # ------------------------------------------------------------------ introspection (:516-568)
    def get_term_stats(self, term: str) -> Optional[Dict]:
        self._require_loaded()
        j = self._term_index.get(term.lower(), -1)
        if j < 0:
            return None
        df = int(self._df_global[j])
        n = self.tables.total_docs
        return {"term": term, "document_frequency": df, "total_frequency": int(self.tables.total_freq[j]),
                "inverse_document_frequency": math.log((n - df + 0.5) / (df + 0.5))}   # natural log, as :532

    def get_document_terms(self, doc_id: int, limit: int = 20) -> List[Tuple[str, int]]:
        self._require_loaded()
        t = self.tables
        pos = int(np.searchsorted(t.doc_ids, doc_id))
        if pos >= len(t.doc_ids) or t.doc_ids[pos] != doc_id:
            return []
        hit = np.flatnonzero(t.post_doc == pos)
        term_of = np.searchsorted(t.term_off, hit, side="right") - 1
        order = np.argsort(-t.post_tf[hit], kind="stable")[:limit]
        return [(t.terms[int(term_of[i])], int(t.post_tf[hit[i]])) for i in order]

    def get_index_stats(self) -> Dict:
        self._require_loaded()
        t = self.tables
        processed = int(np.count_nonzero(t.doc_len))
        total = len(self.store.all_doc_ids()) if hasattr(self.store, "all_doc_ids") else processed
        return {"total_documents_in_database": total, "processed_documents": processed,
                "unique_terms": len(t.terms) if t.terms is not None else int(len(t.term_off) - 1),
                "average_document_length": t.avgdl,
                "index_coverage": f"{processed}/{total} ({100 * processed / max(total, 1):.1f}%)"}

    # ------------------------------------------------------------------ index build (:252-369)
    def build_index(self, batch_size: int = 5000):
        """Build of the four bm25_* tables from ``urlsDB``, then upload.  The text front-end (tokenisation, term
        dictionary) stays on the host as in the reference; the aggregation of the tokens into per-document term
        frequencies, document frequencies and total frequencies runs on the GPU (``mse_bm25_aggregate``,
        SURVEY.md §8f N4); the float32 corpus statistics and the float32 log10 IDF are formed here exactly as
        ``bm25_indexer.py:130-147, 346-369`` do."""
        if not isinstance(self.store, SqlStore):
            raise RuntimeError("build_index needs an SQL store holding urlsDB")
        vocab: Dict[str, int] = {}
        doc_ids: List[int] = []
        tok_off: List[int] = [0]
        tok_term: List[int] = []
        for doc_id, title, text in self.store.iter_documents():
            s = f"{title or ''} {text or ''}".lower().replace("tuebingen", "tübingen").replace("tubingen", "tübingen")[:1_000_000]
            toks = self._tokenize(s)
            if not toks:
                continue
            doc_ids.append(doc_id)
            tok_term.extend(vocab.setdefault(w, len(vocab)) for w in toks)
            tok_off.append(len(tok_term))
        n = len(doc_ids)
        terms = list(vocab)                                   # term id -> string (insertion order)
        term_off, post_doc, post_tf, total_freq = _native.bm25_aggregate(np.asarray(tok_off, np.int64), np.asarray(tok_term, np.int32),
                                                                         len(terms), self.device)
        doc_len = np.diff(np.asarray(tok_off, np.int64))
        df = np.diff(term_off)
        avg = float(np.float32(np.mean(doc_len))) if n else 0.0
        n32 = float(np.float32(n))
        idf_of = lambda d: float(np.float32(math.log10((n32 - d + 0.5) / (d + 0.5))))
        doc_rows = [(doc_ids[i], int(doc_len[i])) for i in range(n)]
        term_of_post = np.repeat(np.arange(len(terms)), df)
        tf_rows = [(doc_ids[int(d)], terms[int(t)], int(f)) for t, d, f in zip(term_of_post, post_doc, post_tf)]
        self.store.write_bm25(doc_rows, tf_rows, [(terms[t], int(df[t]), int(total_freq[t])) for t in range(len(terms))], avg, n, idf_of)
        self.reload()

    def _require_loaded(self):
        if self.native is None or self.tables is None:
            raise RuntimeError("BM25 index is not loaded (no bm25_* tables in the store; call build_index())")

    def close(self):
        if self.native is not None:
            self.native.close()
            self.native = None


def bm25_from_arrays(term_off, post_doc, post_tf, doc_len, idf, avgdl, total_docs, doc_ids=None, terms=None,
                     k1: float = 1.2, b: float = 0.75, device: int = 0, doc_range=None, tokenizer=None) -> BM25:
    """BM25 façade over arrays already in memory (synthetic corpora: integer term ids)."""
    n = len(doc_len)
    ids = np.asarray(doc_ids if doc_ids is not None else np.arange(1, n + 1), dtype=np.int64)
    t = Bm25Tables(terms, np.asarray(term_off, dtype=np.int64), np.asarray(post_doc, dtype=np.int32),
                   np.asarray(post_tf, dtype=np.int32), ids, np.asarray(doc_len, dtype=np.int32),
                   np.asarray(idf, dtype=np.float32), np.zeros(len(term_off) - 1, dtype=np.int64), float(avgdl), float(total_docs))
    return BM25(None, k1, b, store=ArrayStore(bm25=t), tokenizer=tokenizer or whitespace_tokenizer,
                device=device, doc_range=doc_range)
