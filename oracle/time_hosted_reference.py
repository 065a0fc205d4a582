"""TEST INFRASTRUCTURE — times the UNMODIFIED reference (``/root/reference/indexer/bm25_indexer.py::BM25.search`` ->
``/root/reference/reranker/reranker_api.py::rerank``) on the C1 workload of BASELINE.json (configs[0]: 10 000 synthetic docs,
50 000 random 768-d fp32 chunk embeddings, BM25 top-100 + dense rerank), hosted under ``oracle/stub_harness.py``, and the
oracle port on the same queries beside it.  Run in the build container only (the reference is not on the GPU box):

    python oracle/time_hosted_reference.py            # writes profiles/hosted_reference_c1.json

``bench.py`` cites the committed file in its ``cpu_baseline`` record (``hosted_reference_c1``): it is the provenance of the
CPU baseline — the reference's own code and SQL, with sqlite3 standing in for DuckDB, a whitespace tokeniser for spaCy and a
table of query vectors for SentenceTransformer — not a number measured by the run that prints it.
"""
from __future__ import annotations

import json
import os
import platform
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import bm25_oracle as bo  # noqa: E402
from oracle import rerank_oracle as ro  # noqa: E402
from oracle import stub_harness as sh  # noqa: E402

N_DOCS, N_CHUNKS, VOCAB, SEED, TOP_K, N_QUERIES = 10_000, 50_000, 200_000, 1234, 100, 32


def main():
    import torch
    import mse_b200  # noqa: F401
    from mse_b200 import synthetic
    c = synthetic.make_bm25_corpus(N_DOCS, vocab=VOCAB, seed=SEED, always_frac=0.95)
    d = synthetic.make_dense_corpus(N_DOCS, seed=SEED, device="cpu", dtype=torch.float32, total_chunks=N_CHUNKS)
    term_off, post_doc, post_tf = c.term_off.numpy(), c.post_doc.numpy(), c.post_tf.numpy()
    doc_len = c.doc_len.numpy()
    doc_ids = np.arange(1, N_DOCS + 1, dtype=np.int64)               # urlsDB ids, ascending
    names = [f"t{i}" for i in range(c.n_terms)]
    names[c.always_term] = "tübingen"
    q_off, q_term, q_tf = synthetic.make_bm25_queries(c, N_QUERIES, terms_per_query=4, min_rank=64, seed=SEED + 1, add_always=True)
    queries = [" ".join(names[int(t)] for s in range(q_off[i], q_off[i + 1]) for t in [q_term[s]] * int(q_tf[s])) for i in range(N_QUERIES)]
    qv = synthetic.make_query_vectors(N_QUERIES, seed=SEED + 2)
    chunk_off = d.doc_chunk_off.numpy()
    chunk_doc = np.repeat(doc_ids, np.diff(chunk_off))
    emb = d.emb.numpy()

    out = {"workload": f"C1: {N_DOCS} synthetic docs (Zipf(1.0) vocab {VOCAB}, always-term in 95 % of docs appended to every query), "
                       f"{N_CHUNKS} x 768 fp32 chunks, BM25 top-{TOP_K} -> rerank -> top-{TOP_K}, {N_QUERIES} queries one at a time",
           "host": {"cpu_count": os.cpu_count(), "machine": platform.machine(), "python": platform.python_version(),
                    "numpy": np.__version__, "where": "build container (the GPU box has no /root/reference)"},
           "stand_ins": "sqlite3 for DuckDB (LOG = float32 log10, FIRST aggregate), whitespace tokeniser for spaCy, table of query "
                        "vectors for SentenceTransformer, pandas-2 groupby.apply shim (oracle/stub_harness.py)"}
    with sh.hosted_reference() as h:
        sh.create_urls(h.raw, [(int(i), f"https://d{int(i) % 97}.example/{int(i)}", f"title {int(i)}", f"text of document {int(i)}") for i in doc_ids])
        sh.create_chunks(h.raw, np.arange(N_CHUNKS, dtype=np.int64), chunk_doc, emb)
        bm = h.BM25("ignored.db", read_only=False)                    # the reference's own constructor creates the bm25_* tables
        sh.bulk_load_bm25_tables(h.raw, doc_ids, doc_len, names, term_off, doc_ids[post_doc], post_tf)
        bm._update_corpus_stats()                                     # the reference's own statistics and idf statements
        bm._recalculate_idf_scores()
        mod = h.load_reranker()
        mod.config["similarity"]["diversification"] = False
        t_bm = t_rr = 0.0
        n_res = 0
        ref_lists = []
        for i, q in enumerate(queries):
            t0 = time.perf_counter()
            res = bm.search(q, top_k=TOP_K)
            t1 = time.perf_counter()
            resp = h.run_rerank(mod, [r["doc_id"] for r in res], [r["score"] for r in res], q, qv[i])
            t2 = time.perf_counter()
            t_bm += t1 - t0
            t_rr += t2 - t1
            n_res += len(resp.document_scores)
            ref_lists.append(([r["doc_id"] for r in res], [r["score"] for r in res],
                              [int(x.doc_id) for x in resp.document_scores], [x.similarity_score for x in resp.document_scores]))
    out["hosted_reference"] = {"queries_per_s": N_QUERIES / (t_bm + t_rr), "bm25_search_ms_per_query": 1e3 * t_bm / N_QUERIES,
                               "rerank_ms_per_query": 1e3 * t_rr / N_QUERIES, "results_per_query": n_res / N_QUERIES,
                               "path": "BM25.search(top_k=100) -> rerank() of the unmodified reference modules, one query at a time"}

    # the oracle port on the same queries (what bench.py times on the GPU box), checked against the hosted reference
    ix = bo.Bm25Arrays(term_off, post_doc, post_tf, doc_len, c.idf.numpy(), c.avgdl, c.total_docs, doc_ids)
    dense = ro.DenseArrays(emb, np.arange(N_CHUNKS, dtype=np.int64), chunk_off, doc_ids,
                           [f"https://d{int(i) % 97}.example/{int(i)}" for i in doc_ids])
    t_port = {"faithful": 0.0, "fast": 0.0}
    agree = 0
    for faithful in (True, False):
        for i in range(N_QUERIES):
            terms = [int(t) for s in range(q_off[i], q_off[i + 1]) for t in [q_term[s]] * int(q_tf[s])]
            t0 = time.perf_counter()
            b = (bo.search_faithful if faithful else bo.search_fast)(ix, terms, top_k=TOP_K, min_score=0.0)
            r = ro.rerank(dense, np.asarray([x for x, _ in b], dtype=np.int64), np.asarray([s for _, s in b]), qv[i], faithful=faithful)
            t_port["faithful" if faithful else "fast"] += time.perf_counter() - t0
            if faithful:
                rb, rs, rd, rf = ref_lists[i]
                same = [int(doc_ids[x]) for x, _ in b] == rb and np.allclose([s for _, s in b], rs, rtol=1e-6, atol=0) \
                    and [int(doc_ids[x]) for x in r.doc] == rd and np.allclose(r.score, rf, rtol=0, atol=1e-5)
                agree += int(same)
    out["oracle_port"] = {"faithful_queries_per_s": N_QUERIES / t_port["faithful"], "fast_queries_per_s": N_QUERIES / t_port["fast"],
                          "queries_identical_to_the_hosted_reference": agree, "queries": N_QUERIES,
                          "note": "in-memory arrays: no SQL, tokeniser or HTTP cost (a conservative baseline)"}
    path = os.path.join(ROOT, "profiles", "hosted_reference_c1.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, ensure_ascii=False)
    print(json.dumps(out, indent=1, ensure_ascii=False))


if __name__ == "__main__":
    if not sh.reference_available():
        raise SystemExit("the reference is not mounted at " + sh.REFERENCE_ROOT)
    main()
