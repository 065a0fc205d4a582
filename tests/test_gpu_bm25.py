"""GPU parity tests for Stage 1 (BM25 scoring + top-k), all through the C ABI (ctypes).
Oracle = oracle/bm25_oracle.py (pinned against the reference in test_oracle_bm25.py)."""
import json
import os
import sqlite3

import numpy as np
import pytest
import torch

import mse_b200
import mse_testlib as helpers
from mse_b200 import _native, synthetic
from mse_b200.bm25_indexer import BM25, bm25_from_arrays, whitespace_tokenizer
from mse_b200.store import SqlStore
from oracle import bm25_oracle as bo

pytestmark = pytest.mark.gpu

RTOL = 1e-5     # north_star: BM25 scores within 1e-5 relative in fp32


def _oracle_arrays(c):
    return bo.Bm25Arrays(c.term_off.cpu().numpy(), c.post_doc.cpu().numpy(), c.post_tf.cpu().numpy(),
                         c.doc_len.cpu().numpy(), c.idf.cpu().numpy(), c.avgdl, c.total_docs, c.doc_ids.cpu().numpy())


def _facade(ix: bo.Bm25Arrays, **kw):
    return bm25_from_arrays(ix.term_off, ix.post_doc, ix.post_tf, ix.doc_len, ix.idf, ix.avgdl, ix.total_docs,
                            doc_ids=ix.doc_ids, terms=ix.terms, **kw)


def _check_batch(ix, q_off, q_term, q_tf, doc, score, count, top_k, min_score):
    for i in range(len(q_off) - 1):
        terms = []
        for s in range(q_off[i], q_off[i + 1]):
            terms += [int(q_term[s])] * int(q_tf[s])
        ref = bo.search_fast(ix, terms, top_k=top_k, min_score=min_score)
        n = int(count[i])
        scale = bo.abs_contrib_sum(ix, terms)
        ref_doc = [d for d, _ in ref]
        helpers.assert_topk_matches(doc[i, :n], score[i, :n], ref_doc, [s for _, s in ref], RTOL,
                                    scale=scale[ref_doc] if ref_doc else None)
        assert (doc[i, n:] == -1).all()


def test_appendix_e_through_sql_store_and_facade():
    """The reference's own known answers through BM25.search on the GPU."""
    ix, e = helpers.load_appendix_e()
    conn = sqlite3.connect(":memory:")
    conn.execute("CREATE TABLE urlsDB (id BIGINT PRIMARY KEY, url TEXT, title TEXT, text TEXT)")
    conn.executemany("INSERT INTO urlsDB VALUES (?,?,?,?)", [(i, f"http://x/{i}", "", t) for i, t in e["docs"]])
    bm = BM25(None, store=SqlStore(conn), tokenizer=whitespace_tokenizer, load=False)
    bm.build_index()                                   # host build of the four tables + upload
    rows = conn.execute("SELECT term, doc_freq, total_freq, idf_score FROM bm25_term_stats ORDER BY term").fetchall()
    assert [list(r) for r in rows] == e["term_stats"]  # float32 log10 idf, bit-exact with the reference's tables
    for s in e["searches"]:
        got = bm.search(s["query"], top_k=s["top_k"], min_score=s["min_score"])
        assert [g["doc_id"] for g in got] == [r[0] for r in s["result"]], s["query"]
        np.testing.assert_allclose([g["score"] for g in got], [r[1] for r in s["result"]], rtol=RTOL, atol=1e-7)
        assert [g["text_snippet"] for g in got] == [r[2] for r in s["result"]]
    assert bm.get_term_stats("beta")["document_frequency"] == 4
    assert bm.get_index_stats()["processed_documents"] == 6
    bm.close()


@pytest.mark.parametrize("range_docs,qpi,use_tau,accum", [(0, 0, 1, 0), (256, 1, 1, 0), (256, 3, 0, 0), (0, 0, 1, 16), (0, 2, 0, 16)])
def test_golden_small_corpus(range_docs, qpi, use_tau, accum):
    """tests/golden/bm25_small: results of the unmodified reference search().  accum = 16 forces the two-phase kernel
    (bm25_u16.cuh) wherever it applies (min_score >= 0), which a corpus this small would not get by default."""
    ix, j, _ = helpers.load_bm25_small()
    bm = _facade(ix)
    bm.native.set_option("bm25_accum", accum)
    bm.native.set_option("bm25_range_docs", range_docs)
    bm.native.set_option("bm25_queries_per_item", qpi)
    bm.native.set_option("bm25_use_tau", use_tau)
    for s in j["searches"]:
        ids, score, count = bm.search_batch([s["query"]], top_k=min(s["top_k"], 1000), min_score=s["min_score"])
        n = int(count[0])
        terms = s["query"].split()
        scale = bo.abs_contrib_sum(ix, terms)
        ref_idx = np.searchsorted(ix.doc_ids, s["doc_ids"])
        helpers.assert_topk_matches(ids[0, :n], score[0, :n], s["doc_ids"], s["scores"], RTOL,
                                    scale=scale[ref_idx] if len(ref_idx) else None)
    bm.close()


@pytest.fixture(scope="module")
def corpus20k():
    c = synthetic.make_bm25_corpus(20_000, vocab=5_000, mean_len=64, seed=7, always_frac=0.95)
    return c, _oracle_arrays(c)


@pytest.mark.parametrize("opts", [
    dict(),                                                   # defaults
    dict(bm25_range_docs=512, bm25_queries_per_item=1),       # many ranges, 1 query per item
    dict(bm25_range_docs=2048, bm25_queries_per_item=5, bm25_use_tau=0),
    dict(bm25_cand_cap=64),                                   # forces the overflow re-run path
    dict(bm25_range_docs=24576),                              # one range holds the whole corpus
    dict(bm25_accum=16),                                      # two-phase kernel (16-bit bounds, exact rescoring); fp32 kernel at min_score < 0
    dict(bm25_accum=16, bm25_queries_per_item=1, bm25_use_tau=0),   # no running bound: every task ends in its exact mode
    dict(bm25_accum=16, bm25_cand_cap=64),                    # two-phase kernel + the overflow re-run path
])
@pytest.mark.parametrize("top_k,min_score", [(10, 0.0), (1000, 0.0), (100, -50.0)])
def test_synthetic_batch_vs_oracle(corpus20k, opts, top_k, min_score):
    c, ix = corpus20k
    bm = _facade(ix)
    for k, v in opts.items():
        bm.native.set_option(k, v)
    q_off, q_term, q_tf = synthetic.make_bm25_queries(c, 48, terms_per_query=4, min_rank=8, seed=11,
                                                       repeat_frac=0.2, add_always=True)
    doc, score, count = bm.search_batch_terms(q_off, q_term, q_tf, top_k, min_score)
    _check_batch(ix, q_off, q_term, q_tf, doc, score, count, top_k, min_score)
    st = bm.native.bm25_stats()
    # every posting of every query term is either streamed or (negative-idf terms, min_score >= 0) looked up per candidate
    assert st["postings"] + st["postings_looked_up"] == int(np.diff(ix.term_off)[q_term].sum())
    assert (st["postings_looked_up"] > 0) == (min_score >= 0.0)
    if opts.get("bm25_cand_cap") == 64:
        assert st["rerun_queries"] > 0
    bm.close()


@pytest.mark.parametrize("opts", [
    dict(bm25_tau_init=0),                                    # no impact-table seed: the running bound alone
    dict(bm25_readout=0),                                     # candidates by a scan of the accumulators
    dict(bm25_readout=0, bm25_range_docs=512, bm25_queries_per_item=3),
    dict(bm25_range_docs=1536),                               # sub-ranges that are not a power of two
    dict(bm25_range_docs=2048),                               # skip table with two entries per sub-range
    dict(bm25_accum=16),                                      # two-phase kernel: the same fp32 scores bit for bit
    dict(bm25_accum=16, bm25_tau_init=0, bm25_queries_per_item=3),
])
def test_kernel_variants_agree(corpus20k, opts):
    """Every variant returns what the default configuration returns (same summation order)."""
    c, ix = corpus20k
    q_off, q_term, q_tf = synthetic.make_bm25_queries(c, 64, terms_per_query=4, min_rank=8, seed=21,
                                                       repeat_frac=0.2, add_always=True)
    bm = _facade(ix)
    ref = bm.search_batch_terms(q_off, q_term, q_tf, 200, 0.0)
    for k, v in opts.items():
        bm.native.set_option(k, v)
    got = bm.search_batch_terms(q_off, q_term, q_tf, 200, 0.0)
    assert np.array_equal(ref[2], got[2])
    if opts.get("bm25_accum") == 16:                          # same terms, same order, same round-down FMAs
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])
        assert bm.native.bm25_stats()["exact_mode_tasks"] > 0
    np.testing.assert_allclose(got[1], ref[1], rtol=2e-6, atol=4e-6)
    assert np.mean(ref[0] == got[0]) > 0.995
    _check_batch(ix, q_off, q_term, q_tf, got[0], got[1], got[2], 200, 0.0)
    bm.close()


@pytest.mark.parametrize("top_k,use_tau", [(10, 1), (300, 1), (40, 0)])
def test_two_phase_kernel_queue_path_long_queries(top_k, use_tau):
    """The two-phase kernel where its normal path runs: 300k docs = 98 sub-ranges of 3072, small k, so that thousands of tasks
    queue their hits instead of ending in exact mode; queries of 9 terms (more than the four prefetched slots), several
    of them with negative idf (the generic looked-up path of the rescoring pass), one term twice.  Against the fp32
    kernel bit for bit, and against the oracle."""
    c = synthetic.make_bm25_corpus(300_000, vocab=20_000, mean_len=48, seed=5, always_frac=0.95)
    ix = _oracle_arrays(c)
    # (480 queries: 47k tasks = ten waves of warps — in the first wave no query has a bound yet and every task runs in exact mode)
    q_off, q_term, q_tf = synthetic.make_bm25_queries(c, 480, terms_per_query=9, min_rank=2, seed=41, repeat_frac=0.3, add_always=True)
    neg_per_query = np.add.reduceat((np.asarray(ix.idf)[q_term] < 0).astype(np.int64), q_off[:-1])
    assert neg_per_query.max() >= 2                                      # some query holds a second negative-idf term
    bm = _facade(ix)
    bm.native.set_option("bm25_use_tau", use_tau)
    bm.native.set_option("bm25_accum", 32)
    ref = bm.search_batch_terms(q_off, q_term, q_tf, top_k, 0.0)
    assert bm.native.bm25_stats()["exact_mode_tasks"] == 0
    bm.native.set_option("bm25_accum", 16)
    got = bm.search_batch_terms(q_off, q_term, q_tf, top_k, 0.0)
    st = bm.native.bm25_stats()
    assert st["ranges"] == -(-300_000 // 3072)
    if use_tau:
        assert 0 < st["exact_mode_tasks"] <= 480 * st["ranges"]
        if top_k == 10:
            assert st["exact_mode_tasks"] < 0.95 * 480 * st["ranges"]    # the queue path took the rest (measured: a third of the tasks)
    for r, g in zip(ref, got):
        assert np.array_equal(r, g)
    _check_batch(ix, q_off[:9], q_term, q_tf, got[0], got[1], got[2], top_k, 0.0)
    bm.close()


def test_impact_table_matches_checker(corpus20k):
    """The seed of the candidate filter: GPU results with and without it are identical, and the bound the oracle-side
    restatement derives for these queries is below the k-th returned score."""
    c, ix = corpus20k
    levels = bo.impact_levels(ix)
    q_off, q_term, q_tf = synthetic.make_bm25_queries(c, 32, terms_per_query=3, min_rank=8, seed=31, add_always=False)
    bm = _facade(ix)
    doc, score, count = bm.search_batch_terms(q_off, q_term, q_tf, 100, 0.0)
    for i in range(32):
        terms = [int(t) for s in range(q_off[i], q_off[i + 1]) for t in [q_term[s]] * int(q_tf[s])]
        bound = bo.initial_bound(ix, levels, terms, 100)
        if bound > 0:
            assert count[i] == 100 and score[i, 99] >= bound * (1 - 1e-6)
    bm.close()


def test_edge_cases(corpus20k):
    c, ix = corpus20k
    bm = _facade(ix)
    V = ix.n_terms
    # q0: no terms; q1: only unknown / out-of-range terms; q2: rare term (fewer docs than k); q3: always-term only
    # (negative idf -> everything < 0 -> empty at min_score 0); q4: same with min_score admitted below
    df = np.diff(ix.term_off)
    rare = int(np.flatnonzero((df > 0) & (df < 40))[0])
    q_off = np.asarray([0, 0, 2, 3, 4, 5], dtype=np.int32)
    q_term = np.asarray([-5, V + 10, rare, c.always_term, c.always_term], dtype=np.int32)
    q_tf = np.ones(5, dtype=np.int32)
    doc, score, count = bm.search_batch_terms(q_off, q_term, q_tf, 50, 0.0)
    assert count.tolist()[:3] == [0, 0, int(df[rare])] and count[3] == 0
    assert float(ix.idf[c.always_term]) < 0
    doc2, score2, count2 = bm.search_batch_terms(q_off, q_term, q_tf, 50, -100.0)
    assert count2[4] == 50 and np.all(np.diff(score2[4, :50]) <= 0) and score2[4, 0] < 0
    _check_batch(ix, q_off, q_term, q_tf, doc2, score2, count2, 50, -100.0)
    # errors surface as exceptions, never as silent fallbacks
    with pytest.raises(_native.NativeError):
        bm.search_batch_terms(q_off, q_term, q_tf, _native.MAX_TOPK + 1, 0.0)
    bm.close()


def test_zero_idf_docs_are_kept_and_tie_broken_by_doc():
    """Appendix E 'alpha': idf == 0 -> touched docs with score exactly 0.0 are returned in ascending doc id."""
    ix, _ = helpers.load_appendix_e()
    bm = _facade(ix)
    ids, score, count = bm.search_batch(["alpha"], top_k=10)
    assert ids[0, :count[0]].tolist() == [1, 3, 6] and np.all(score[0, :3] == 0.0)
    ids, score, count = bm.search_batch(["alpha", "beta", "nonexistent"], top_k=2)
    assert ids[0, :2].tolist() == [1, 3] and count.tolist() == [2, 0, 0]
    bm.close()


def test_device_resident_path_equals_host_path(corpus20k):
    c, ix = corpus20k
    bm = _facade(ix)
    q_off, q_term, q_tf = synthetic.make_bm25_queries(c, 32, min_rank=8, seed=3)
    h = bm.search_batch_terms(q_off, q_term, q_tf, 100, 0.0)
    dev = torch.device("cuda:0")
    d = bm.search_batch_terms(torch.from_numpy(q_off).to(dev), torch.from_numpy(q_term).to(dev),
                              torch.from_numpy(q_tf).to(dev), 100, 0.0)
    torch.cuda.synchronize()
    for a, b in zip(h, d):
        np.testing.assert_array_equal(a, b.cpu().numpy())     # deterministic: no float atomics on the path
    bm.close()


def test_malformed_index_is_rejected():
    nat = _native.NativeIndex(0)
    term_off = np.asarray([0, 2], dtype=np.int64)
    with pytest.raises(_native.NativeError):      # descending docs inside a term
        nat.bm25_load(term_off, np.asarray([3, 1], np.int32), np.asarray([1, 1], np.int32),
                      np.asarray([5, 5, 5, 5], np.int32), np.asarray([0.5], np.float32), 5.0)
    with pytest.raises(_native.NativeError):      # search before load
        nat.bm25_search(np.asarray([0, 1], np.int32), np.asarray([0], np.int32), np.asarray([1], np.int32), 10)
    nat.close()


def test_topk_merge_matches_host_rule():
    from mse_b200.sharding import merge_topk_host
    rng = np.random.default_rng(0)
    W, B, k = 4, 9, 37
    g_doc = np.full((W, B, k), -1, np.int32); g_score = np.zeros((W, B, k), np.float32)
    g_count = rng.integers(0, k + 1, size=(W, B)).astype(np.int32)
    for w in range(W):
        for q in range(B):
            n = g_count[w, q]
            s = np.sort(rng.integers(0, 12, size=n).astype(np.float32) / 4)[::-1]      # many ties across lists
            g_doc[w, q, :n] = w * 1000 + np.sort(rng.choice(1000, size=n, replace=False))
            g_score[w, q, :n] = s
    nat = _native.NativeIndex(0)
    for top_k in (5, 64):
        got = nat.topk_merge(g_doc, g_score, g_count, top_k)
        want = merge_topk_host(g_doc, g_score, g_count, top_k)
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)
    nat.close()


def test_sharded_equals_unsharded(corpus20k):
    """Doc-range shards scored separately + merge == one index (the single-process view of §8e)."""
    from mse_b200.bm25_indexer import shard_bounds
    c, ix = corpus20k
    q_off, q_term, q_tf = synthetic.make_bm25_queries(c, 24, min_rank=8, seed=5, add_always=True)
    whole = _facade(ix)
    ref = whole.search_batch_terms(q_off, q_term, q_tf, 200, 0.0)
    per_doc = np.bincount(ix.post_doc, minlength=ix.n_docs)
    bounds = shard_bounds(per_doc, 3)
    parts = []
    for r in range(3):
        sh = _facade(ix, doc_range=(bounds[r], bounds[r + 1]))
        parts.append(sh.search_batch_terms(q_off, q_term, q_tf, 200, 0.0))
        sh.close()
    g = [np.stack([p[i] for p in parts]) for i in range(3)]
    merged = whole.native.topk_merge(g[0], g[1], g[2], 200)
    for a, b in zip(ref, merged):
        np.testing.assert_array_equal(a, b)
    whole.close()


def test_shard_lists_cut_to_m_merge_exactly_when_the_rule_says_so(corpus20k):
    """Query-owner exchange (sharding.exchange_topk_owner) without the collective: shard-local searches with
    top_k = m < k, merged; whenever ``cut_could_hide_a_result`` reports the cut safe the merge equals the unsharded
    top-k bit for bit, and m == k is always safe."""
    from mse_b200.bm25_indexer import shard_bounds
    from mse_b200.sharding import cut_could_hide_a_result
    c, ix = corpus20k
    K, W = 200, 4
    q_off, q_term, q_tf = synthetic.make_bm25_queries(c, 32, min_rank=8, seed=15, add_always=True)
    whole = _facade(ix)
    ref = whole.search_batch_terms(q_off, q_term, q_tf, K, 0.0)
    bounds = shard_bounds(np.bincount(ix.post_doc, minlength=ix.n_docs), W)
    shards = [_facade(ix, doc_range=(bounds[r], bounds[r + 1])) for r in range(W)]
    verdicts = {}
    for m in (K, 2 * K // W + 32, K // W + 8, 5):
        parts = [sh.search_batch_terms(q_off, q_term, q_tf, m, 0.0) for sh in shards]
        g = [np.stack([p[i] for p in parts]) for i in range(3)]
        merged = whole.native.topk_merge(g[0], g[1], g[2], K)
        unsafe = bool(cut_could_hide_a_result(g[1], g[2], merged[1], merged[2], K)) if m < K else False
        verdicts[m] = unsafe
        if not unsafe:
            for a, b in zip(ref, merged):
                np.testing.assert_array_equal(a, b)
    assert verdicts[K] is False and verdicts[2 * K // W + 32] is False and verdicts[5] is True
    for sh in shards:
        sh.close()
    whole.close()


@pytest.mark.parametrize("n_docs", [1_000_000])
def test_full_size_c2_properties_and_sampled_parity(n_docs):
    """BASELINE.json configs[1]: 1M docs, Zipf vocab 200k, batch 1024, top-1000.  Checks
    size-independent properties on the whole batch and exact parity on a sample of queries."""
    dev = torch.device("cuda:0")
    c = synthetic.make_bm25_corpus(n_docs, vocab=200_000, seed=1234, device=dev)
    nat = _native.NativeIndex(0)
    nat.bm25_load(c.term_off, c.post_doc, c.post_tf, c.doc_len, c.idf, c.avgdl)
    q_off, q_term, q_tf = synthetic.make_bm25_queries(c, 1024, seed=1235)
    doc, score, count = nat.bm25_search(q_off, q_term, q_tf, 1000, 0.0)
    st = nat.bm25_stats()
    ix = _oracle_arrays(c)
    df = np.diff(ix.term_off)
    assert st["postings"] + st["postings_looked_up"] == int(df[q_term].sum())
    for i in range(1024):
        n = int(count[i])
        assert np.all(np.diff(score[i, :n]) <= 0)                              # sorted
        assert len(set(doc[i, :n].tolist())) == n                              # unique docs
        ties = np.flatnonzero(np.diff(score[i, :n]) == 0)
        assert np.all(doc[i, ties] < doc[i, ties + 1])                         # ties -> ascending doc
    sample = np.random.default_rng(1).choice(1024, size=12, replace=False)
    sub_off = np.zeros(len(sample) + 1, np.int32)
    for j, i in enumerate(sample):
        sub_off[j + 1] = sub_off[j] + (q_off[i + 1] - q_off[i])
    sub_term = np.concatenate([q_term[q_off[i]:q_off[i + 1]] for i in sample])
    sub_tf = np.concatenate([q_tf[q_off[i]:q_off[i + 1]] for i in sample])
    _check_batch(ix, sub_off, sub_term, sub_tf, doc[sample], score[sample], count[sample], 1000, 0.0)
    # idempotence / batch-composition independence: the sample alone gives the same answer
    d2, s2, c2 = nat.bm25_search(sub_off, sub_term, sub_tf, 1000, 0.0)
    np.testing.assert_array_equal(d2, doc[sample]); np.testing.assert_array_equal(s2, score[sample])
    nat.close()


def test_device_index_aggregation_matches_host_statement():
    """N4: mse_bm25_aggregate (tokens -> CSR postings, df, total_freq) against a numpy statement of
    bm25_indexer.py:181-211 / 283-343, and the façade's build_index() against the oracle's host build."""
    rng = np.random.default_rng(3)
    n_docs, V = 3000, 700
    lens = rng.integers(0, 90, n_docs)
    lens[::97] = 0                                              # empty documents
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tok = np.minimum((rng.pareto(1.1, int(off[-1])) * 3).astype(np.int64), V - 1).astype(np.int32)
    term_off, post_doc, post_tf, total_freq = _native.bm25_aggregate(off, tok, V)
    doc_of = np.repeat(np.arange(n_docs), lens)
    key, cnt = np.unique(tok.astype(np.int64) * n_docs + doc_of, return_counts=True)
    assert np.array_equal(post_doc, (key % n_docs).astype(np.int32)) and np.array_equal(post_tf, cnt.astype(np.int32))
    df = np.bincount((key // n_docs).astype(np.int64), minlength=V)
    assert np.array_equal(np.diff(term_off), df) and term_off[0] == 0
    assert np.array_equal(total_freq, np.bincount(tok, minlength=V))
    with pytest.raises(_native.NativeError):                   # term id outside the dictionary
        _native.bm25_aggregate(off, np.where(np.arange(len(tok)) == 5, V + 3, tok).astype(np.int32), V)
    e0, o0, f0, t0 = _native.bm25_aggregate(np.zeros(1, np.int64), np.zeros(0, np.int32), 4)      # nothing to aggregate
    assert e0.tolist() == [0, 0, 0, 0, 0] and len(o0) == 0 and t0.tolist() == [0, 0, 0, 0]
