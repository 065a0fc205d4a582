"""TEST INFRASTRUCTURE — generates ``tests/golden/*`` by running the UNMODIFIED reference
(``/root/reference/indexer/bm25_indexer.py`` and ``/root/reference/reranker/reranker_api.py``)
under ``oracle/stub_harness.py``.  Run in the build container only:

    python oracle/make_golden.py

Outputs (committed):
  tests/golden/appendix_e.json      the six-document known-answer corpus of SURVEY.md Appendix E
  tests/golden/bm25_small.npz/json  seeded 300-doc corpus: tables as left by the reference's own
                                    build_index, plus reference search results for 24 queries
  tests/golden/rerank_small.npz/json  64-doc chunk/embedding tables and the reference rerank()
                                    responses (with and without diversification)
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import stub_harness as sh  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

APPENDIX_E_DOCS = [
    (1, "alpha beta beta gamma"), (2, "beta gamma gamma gamma delta"), (3, "alpha alpha alpha"),
    (4, "delta epsilon"), (5, "beta"), (6, "beta alpha zeta eta theta iota"),
]
APPENDIX_E_QUERIES = [
    ("gamma delta", 10, 0.0), ("epsilon zeta", 10, 0.0), ("beta gamma", 10, 0.0), ("beta gamma", 10, -1.0),
    ("beta beta delta", 10, 0.0), ("alpha", 10, 0.0), ("beta", 10, 0.0), ("nonexistent", 10, 0.0),
    ("gamma delta", 1, 0.0),
]


def dump_bm25_tables(raw):
    docs = raw.execute("SELECT doc_id, doc_length FROM bm25_doc_stats ORDER BY doc_id").fetchall()
    terms = raw.execute("SELECT term, doc_freq, total_freq, idf_score FROM bm25_term_stats ORDER BY term").fetchall()
    tf = raw.execute("SELECT term, doc_id, freq FROM bm25_term_freq ORDER BY term, doc_id").fetchall()
    stats = dict(raw.execute("SELECT stat_name, stat_value FROM bm25_corpus_stats").fetchall())
    return docs, terms, tf, stats


def run_appendix_e():
    with sh.hosted_reference() as h:
        sh.create_urls(h.raw, [(i, f"http://x/{i}", "", t) for i, t in APPENDIX_E_DOCS])
        bm = h.BM25("ignored.db", read_only=False)
        bm.build_index(batch_size=4)
        docs, terms, tf, stats = dump_bm25_tables(h.raw)
        out = {"docs": APPENDIX_E_DOCS, "doc_stats": docs, "term_stats": terms, "term_freq": tf,
               "corpus_stats": stats, "searches": []}
        for q, k, ms in APPENDIX_E_QUERIES:
            res = bm.search(q, top_k=k, min_score=ms)
            out["searches"].append({"query": q, "top_k": k, "min_score": ms,
                                    "result": [[r["doc_id"], r["score"], r["text_snippet"]] for r in res]})
    with open(os.path.join(GOLD, "appendix_e.json"), "w") as f:
        json.dump(out, f, indent=1, ensure_ascii=False)
    return out


def make_small_docs(n_docs=300, vocab=400, seed=20261018):
    rng = np.random.Generator(np.random.Philox(seed))
    words = [f"w{j:03d}" for j in range(vocab)]
    p = 1.0 / np.arange(1, vocab + 1)
    p /= p.sum()
    docs = []
    for i in range(n_docs):
        L = int(max(3, round(rng.lognormal(np.log(40) - 0.125, 0.5))))
        toks = [words[j] for j in rng.choice(vocab, size=L, p=p)]
        if rng.random() < 0.95:
            toks += ["tübingen"] * int(1 + rng.poisson(3))
        rng.shuffle(toks)
        title = " ".join(toks[:3]) if i % 7 else ""
        doc_id = 1 + 3 * i + int(rng.integers(0, 3))        # non-contiguous ids
        docs.append((doc_id, title, " ".join(toks[3:]) if i % 7 else " ".join(toks)))
    return docs, words


def run_bm25_small():
    docs, words = make_small_docs()
    rng = np.random.Generator(np.random.Philox(99))
    queries = []
    for qi in range(24):
        n = int(rng.integers(1, 5))
        # mix of frequent and rare words; some repeated; some unknown; most with the always-term
        ts = [words[int(min(399, rng.zipf(1.3) + rng.integers(0, 30)))] for _ in range(n)]
        if qi % 5 == 0:
            ts.append(ts[0])
        if qi % 6 == 1:
            ts.append("zzzunknown")
        if qi % 4 != 3:
            ts.append("tübingen")
        queries.append((" ".join(ts), [50, 1000, 7][qi % 3], [0.0, -100.0, 0.05][qi % 3 if qi % 2 else 0]))
    with sh.hosted_reference() as h:
        sh.create_urls(h.raw, [(d, f"https://d{d % 13}.example/{d}", t, x) for d, t, x in docs])
        bm = h.BM25("ignored.db", read_only=False)
        bm.build_index(batch_size=40)       # < 50 docs per batch: sequential path, < sqlite's parameter limit
        d, t, tf, stats = dump_bm25_tables(h.raw)
        results = []
        for q, k, ms in queries:
            res = bm.search(q, top_k=k, min_score=ms)
            results.append({"query": q, "top_k": k, "min_score": ms,
                            "doc_ids": [r["doc_id"] for r in res], "scores": [r["score"] for r in res],
                            "snippet0": res[0]["text_snippet"] if res else None})
    term_names = [r[0] for r in t]
    tix = {n: i for i, n in enumerate(term_names)}
    np.savez_compressed(
        os.path.join(GOLD, "bm25_small.npz"),
        doc_ids=np.asarray([r[0] for r in d], dtype=np.int64),
        doc_len=np.asarray([r[1] for r in d], dtype=np.int32),
        term_df=np.asarray([r[1] for r in t], dtype=np.int64),
        term_total=np.asarray([r[2] for r in t], dtype=np.int64),
        term_idf=np.asarray([r[3] for r in t], dtype=np.float64),   # float32 values held in double
        tf_term=np.asarray([tix[r[0]] for r in tf], dtype=np.int32),
        tf_doc=np.asarray([r[1] for r in tf], dtype=np.int64),
        tf_freq=np.asarray([r[2] for r in tf], dtype=np.int32),
    )
    with open(os.path.join(GOLD, "bm25_small.json"), "w") as f:
        json.dump({"terms": term_names, "corpus_stats": stats, "searches": results,
                   "docs": [[a, b, c] for a, b, c in docs]}, f, ensure_ascii=False)
    return len(d), len(t), len(tf)


def run_rerank_small():
    rng = np.random.Generator(np.random.Philox(4242))
    n_docs, dim = 64, 768
    ids = np.arange(1, n_docs + 1, dtype=np.int64) * 2
    counts = np.minimum(rng.geometric(0.25, n_docs), 14)
    counts[5] = 12; counts[9] = 1; counts[11] = 0
    urls = [f"https://d{int(d) % 9}.example/{int(d)}" for d in ids]
    urls[20] = urls[19] + "?q=x"; urls[33] = urls[30] + "?page=2"   # duplicates collapse to MIN(id)
    chunk_doc = np.repeat(ids, counts)
    chunk_ids = np.arange(len(chunk_doc), dtype=np.int64)
    emb = rng.standard_normal((len(chunk_doc), dim)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    cases = []
    with sh.hosted_reference() as h:
        sh.create_urls(h.raw, [(int(d), u, f"title {int(d)}", f"text of {int(d)}") for d, u in zip(ids, urls)])
        sh.create_chunks(h.raw, chunk_ids, chunk_doc, emb)
        mod = h.load_reranker()
        for ci in range(6):
            k = [64, 40, 10, 3, 1, 25][ci]
            cand = rng.permutation(n_docs)[:k]
            if ci == 0:
                cand = np.arange(n_docs)
            cand_ids = ids[cand]
            sims = np.sort(rng.gamma(2.0, 1.5, size=k))[::-1]
            if ci == 3:
                sims[:] = 1.25                                  # all-equal BM25 -> old == 0
            q = rng.standard_normal(dim).astype(np.float32) * 3.0   # un-normalised (reranker_api.py:355)
            if ci == 1:
                q = (emb[chunk_doc == cand_ids[0]][0] * 2 + 0.05 * q).astype(np.float32)  # a score >= 0.8
            case = {"cand_ids": cand_ids.tolist(), "sims": sims.tolist(), "q_index": ci}
            for div in (False, True):
                mod.config["similarity"]["diversification"] = div
                resp = h.run_rerank(mod, cand_ids, sims, f"query {ci}", q)
                case["div" if div else "plain"] = {
                    "doc_id": [int(x.doc_id) for x in resp.document_scores],
                    "score": [x.similarity_score for x in resp.document_scores],
                    "orig": [x.original_similarity for x in resp.document_scores],
                    "window": [int(x.most_relevant_window.window_index) for x in resp.document_scores],
                    "url": [x.url for x in resp.document_scores],
                    "total_documents": resp.total_documents, "total_windows": resp.total_windows,
                }
            case["q"] = q.tolist()
            cases.append(case)
    np.savez_compressed(os.path.join(GOLD, "rerank_small.npz"), doc_ids=ids, counts=counts,
                        chunk_ids=chunk_ids, chunk_doc=chunk_doc, emb=emb)
    with open(os.path.join(GOLD, "rerank_small.json"), "w") as f:
        json.dump({"urls": urls, "cases": cases}, f)
    return len(chunk_doc)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    e = run_appendix_e()
    print("appendix E:", e["corpus_stats"], [len(s["result"]) for s in e["searches"]])
    print("bm25_small (docs, terms, postings):", run_bm25_small())
    print("rerank_small rows:", run_rerank_small())
