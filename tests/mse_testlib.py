"""Shared test helpers: golden loaders and tolerance-aware top-k comparison."""
import json
import os

import numpy as np

from oracle import bm25_oracle as bo
from oracle import rerank_oracle as ro

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def arrays_from_tables(doc_ids, doc_len, term_names, term_idf, tf_term, tf_doc, tf_freq, avgdl, total_docs):
    """CSR-by-term arrays from rows ordered by (term, doc_id) — what the loader reads out of
    bm25_term_freq / bm25_doc_stats / bm25_term_stats / bm25_corpus_stats."""
    doc_ids = np.asarray(doc_ids, dtype=np.int64)
    order = np.argsort(doc_ids, kind="stable")
    doc_ids = doc_ids[order]
    doc_len = np.asarray(doc_len, dtype=np.int32)[order]
    V = len(term_names)
    tf_term = np.asarray(tf_term)
    term_off = np.zeros(V + 1, dtype=np.int64)
    np.add.at(term_off, tf_term + 1, 1)
    term_off = np.cumsum(term_off)
    post_doc = np.searchsorted(doc_ids, np.asarray(tf_doc, dtype=np.int64)).astype(np.int32)
    return bo.Bm25Arrays(term_off, post_doc, np.asarray(tf_freq, dtype=np.int32), doc_len,
                         np.asarray(term_idf, dtype=np.float32), float(avgdl), float(total_docs), doc_ids,
                         list(term_names), {t: i for i, t in enumerate(term_names)})


def load_bm25_small():
    z = np.load(os.path.join(GOLD, "bm25_small.npz"))
    j = json.load(open(os.path.join(GOLD, "bm25_small.json")))
    ix = arrays_from_tables(z["doc_ids"], z["doc_len"], j["terms"], z["term_idf"], z["tf_term"], z["tf_doc"],
                            z["tf_freq"], j["corpus_stats"]["avg_doc_length"], j["corpus_stats"]["total_docs"])
    return ix, j, z


def load_appendix_e():
    e = json.load(open(os.path.join(GOLD, "appendix_e.json")))
    names = [r[0] for r in e["term_stats"]]
    tix = {t: i for i, t in enumerate(names)}
    ix = arrays_from_tables([r[0] for r in e["doc_stats"]], [r[1] for r in e["doc_stats"]], names,
                            [r[3] for r in e["term_stats"]], [tix[r[0]] for r in e["term_freq"]],
                            [r[1] for r in e["term_freq"]], [r[2] for r in e["term_freq"]],
                            e["corpus_stats"]["avg_doc_length"], e["corpus_stats"]["total_docs"])
    return ix, e


def load_rerank_small():
    z = np.load(os.path.join(GOLD, "rerank_small.npz"))
    j = json.load(open(os.path.join(GOLD, "rerank_small.json")))
    ids = z["doc_ids"]
    off = np.zeros(len(ids) + 1, dtype=np.int64)
    off[1:] = np.cumsum(z["counts"])
    dense = ro.DenseArrays(z["emb"], z["chunk_ids"], off, ids, j["urls"])
    return dense, j


def assert_topk_matches(got_doc, got_score, ref_doc, ref_score, rtol, scale=None, atol=0.0):
    """north_star parity rule: scores within rtol (relative to |score|, or to `scale[doc]` when
    given — the Σ|contribution| floor of SURVEY.md A.1), ids identical except where the scores
    that decide the order are tied within that tolerance."""
    got_doc = np.asarray(got_doc); ref_doc = np.asarray(ref_doc)
    got_score = np.asarray(got_score, dtype=np.float64); ref_score = np.asarray(ref_score, dtype=np.float64)
    assert len(got_doc) == len(ref_doc), (len(got_doc), len(ref_doc))
    if len(ref_doc) == 0:
        return
    tol = rtol * np.maximum(np.abs(ref_score), 0 if scale is None else np.asarray(scale)) + atol
    np.testing.assert_array_less(np.abs(got_score - ref_score), tol + 1e-300)
    mism = np.flatnonzero(got_doc != ref_doc)
    for i in mism:
        # a different doc at rank i is acceptable only if its score ties with the oracle's rank-i score
        assert abs(got_score[i] - ref_score[i]) <= tol[i], (i, got_doc[i], ref_doc[i], got_score[i], ref_score[i])
    if len(mism):
        # and the multisets may differ only in docs sitting at the cut-off score
        diff = set(got_doc.tolist()) ^ set(ref_doc.tolist())
        cut = ref_score[-1]
        for d in diff:
            s = got_score[got_doc == d] if d in set(got_doc.tolist()) else ref_score[ref_doc == d]
            assert abs(s[0] - cut) <= tol[-1] * 2 + atol, (d, s, cut)
