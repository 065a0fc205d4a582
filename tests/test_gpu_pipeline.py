"""End-to-end batch path: queries.txt -> BM25 top-k -> gathered rerank -> fusion ->
diversification -> TSV lines, through an SQL store with the reference's table layout, compared with the
oracle pipeline (bm25 oracle -> rerank oracle -> diversify)."""
import sqlite3

import numpy as np
import pytest
import torch

import mse_b200  # noqa: F401
from mse_b200 import pipeline, synthetic
from mse_b200.bm25_indexer import BM25, whitespace_tokenizer
from mse_b200.reranker import Reranker
from mse_b200.store import SqlStore
from oracle import bm25_oracle as bo
from oracle import rerank_oracle as ro

pytestmark = pytest.mark.gpu


def _make_db(n_docs=400, vocab=300, seed=3):
    rng = np.random.default_rng(seed)
    words = [f"w{j:03d}" for j in range(vocab)]
    p = 1.0 / np.arange(1, vocab + 1)
    p /= p.sum()
    conn = sqlite3.connect(":memory:")
    conn.execute("CREATE TABLE urlsDB (id BIGINT PRIMARY KEY, url TEXT, title TEXT, text TEXT)")
    conn.execute("CREATE TABLE chunks_optimized(chunk_id BIGINT PRIMARY KEY, doc_id BIGINT, chunk_text TEXT)")
    conn.execute("CREATE TABLE embeddings(chunk_id BIGINT PRIMARY KEY, embedding BLOB)")
    ids = np.arange(1, n_docs + 1) * 2 + 1
    urls = synthetic.make_urls(ids, n_domains=23, dup_frac=0.05, seed=seed)
    docs = []
    for d, u in zip(ids.tolist(), urls):
        toks = [words[j] for j in rng.choice(vocab, size=int(rng.integers(5, 60)), p=p)]
        if rng.random() < 0.9:
            toks += ["tübingen"] * int(1 + rng.poisson(2))
        docs.append((d, u, f"title {d}", " ".join(toks)))
    conn.executemany("INSERT INTO urlsDB VALUES (?,?,?,?)", docs)
    counts = synthetic.make_chunk_counts(n_docs, seed=seed)
    counts[7] = 0
    chunk_doc = np.repeat(ids, counts)
    emb = rng.standard_normal((len(chunk_doc), 768)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    conn.executemany("INSERT INTO chunks_optimized VALUES (?,?,?)", [(i, int(d), "c") for i, d in enumerate(chunk_doc)])
    conn.executemany("INSERT INTO embeddings VALUES (?,?)", [(i, e.tobytes()) for i, e in enumerate(emb)])
    return conn, ids, urls, docs, counts, emb


def test_batch_search_file_matches_oracle_pipeline(tmp_path):
    conn, ids, urls, docs, counts, emb = _make_db()
    store = SqlStore(conn)
    bm = BM25(None, store=store, tokenizer=whitespace_tokenizer, load=False)
    bm.build_index()
    rr = Reranker(store, bm.global_doc_ids, native=bm.native, diversification=True)
    queries = ["w003 w010", "w001 tuebingen w020", "w050 w002 w002", "zzz"]
    qfile = tmp_path / "queries.txt"
    qfile.write_text("".join(f"{i + 1}\t{q}\n" for i, q in enumerate(queries)), encoding="utf-8")
    qv = synthetic.make_query_vectors(len(queries), seed=5) * 2.0
    hs = pipeline.HybridSearch(bm, rr, top_k_retrieval=100)
    lines = hs.batch_search_file(str(qfile), str(tmp_path / "out.txt"), query_vecs=qv)
    assert (tmp_path / "out.txt").read_text(encoding="utf-8").splitlines() == lines

    # oracle pipeline on the same tables
    toks = [{w: t.lower().split().count(w) for w in set(t.lower().split())} for t in (f"{d[2]} {d[3]}" for d in docs)]
    ix = bo.build_arrays([d[0] for d in docs], toks)
    off = np.concatenate([[0], np.cumsum(counts)])
    stored = torch.from_numpy(emb).to(torch.bfloat16).float().numpy()
    dense = ro.DenseArrays(stored, np.arange(len(emb)), off, np.asarray([d[0] for d in docs]), urls)
    got = {}
    for ln in lines:
        num, rank, url, score = ln.split("\t")
        got.setdefault(int(num), []).append((int(rank), url, float(score)))
    for i, q in enumerate(queries):
        terms = pipeline.preprocess_query(q).split()
        cand = bo.search_fast(ix, terms, top_k=100, min_score=0.0)
        if not cand:
            assert (i + 1) not in got
            continue
        res = ro.rerank(dense, [d for d, _ in cand], [s for _, s in cand], qv[i], faithful=False)
        sel, sc = ro.diversify([urls[d] for d in res.doc], res.score, 0.8, 100)
        want_urls = [urls[res.doc[j]] for j in sel]
        g = got[i + 1]
        assert [r for r, _, _ in g] == list(range(1, len(g) + 1)) and len(g) == len(want_urls)
        np.testing.assert_allclose([s for _, _, s in g], np.round(sc, 3), atol=3e-3)
        same = sum(a == b for (_, a, _), b in zip(g, want_urls))
        assert same >= 0.9 * len(want_urls)          # order may differ only inside the dense tolerance
    bm.close()


def test_c1_config_against_oracle():
    """BASELINE.json configs[0] shape: 10k synthetic docs (with the always-present 'tübingen' term), 50k
    768-d chunks, the five queries.txt queries mapped to term ids + the always-term, BM25 top-100 -> rerank ->
    top-100.  GPU pipeline (array store) vs the oracle pipeline on the same arrays."""
    from mse_b200.bm25_indexer import bm25_from_arrays
    from mse_b200.store import ArrayStore, DenseTables
    c = synthetic.make_bm25_corpus(10_000, vocab=200_000, seed=1234, always_frac=0.95)
    ix = bo.Bm25Arrays(c.term_off.numpy(), c.post_doc.numpy(), c.post_tf.numpy(), c.doc_len.numpy(), c.idf.numpy(),
                       c.avgdl, c.total_docs, c.doc_ids.numpy())
    d = synthetic.make_dense_corpus(10_000, seed=1234, device="cpu", dtype=torch.float32, total_chunks=50_000)
    emb, off = d.emb.numpy(), d.doc_chunk_off.numpy()
    assert emb.shape[0] == 50_000
    urls = synthetic.make_urls(ix.doc_ids, n_domains=997, dup_frac=0.02)
    bm = bm25_from_arrays(ix.term_off, ix.post_doc, ix.post_tf, ix.doc_len, ix.idf, ix.avgdl, ix.total_docs, doc_ids=ix.doc_ids)
    store = ArrayStore(bm25=bm.tables, doc_id_array=ix.doc_ids, url_list=urls)
    rr = Reranker(store, ix.doc_ids, native=bm.native, dense_tables=DenseTables(emb, d.chunk_ids.numpy(), off), diversification=False)
    stored = torch.from_numpy(emb).to(torch.bfloat16).float().numpy()
    dense = ro.DenseArrays(stored, d.chunk_ids.numpy(), off, ix.doc_ids, urls)
    # queries.txt lines -> 2-4 hashed term ranks (>= 64, with postings) + the always-present term
    df = np.diff(ix.term_off)
    lines = ["tübingen attractions", "food and drinks", "university tuebingen research programs",
             "castle hohentubingen history", "botanical garden opening hours"]
    known = np.flatnonzero(df[:c.always_term] > 0)
    known = known[known >= 64]
    qv = synthetic.make_query_vectors(len(lines), seed=8) * 1.7
    for li, line in enumerate(lines):
        words = [w for w in pipeline.preprocess_query(line).split() if w != "tübingen"]
        import zlib
        terms = [int(known[zlib.crc32(w.encode()) % len(known)]) for w in words] + [c.always_term]
        q_off, q_term, q_tf = bm.encode_queries([terms])
        doc, score, count = bm.search_batch_terms(q_off, q_term, q_tf, 100, 0.0)
        ref = bo.search_fast(ix, terms, top_k=100, min_score=0.0)
        n = int(count[0])
        assert n == len(ref)
        if n == 0:
            continue
        import mse_testlib as helpers
        scale = bo.abs_contrib_sum(ix, terms)
        rd = [x for x, _ in ref]
        helpers.assert_topk_matches(doc[0, :n], score[0, :n], rd, [s for _, s in ref], 1e-5, scale=scale[rd])
        out = rr.rerank_batch([doc[0, :n]], [score[0, :n]], qv[li:li + 1])
        res = ro.rerank(dense, rd, [s for _, s in ref], qv[li], faithful=False)
        m = int(out[4][0])
        assert m == len(res.doc) and int(out[5][0]) == res.total_rows
        helpers.assert_topk_matches(out[0][0, :m], out[1][0, :m], res.doc, res.score, 0.0, atol=2e-3)
    bm.close()
