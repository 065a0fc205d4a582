"""Parity AT THE SIZES BASELINE.json names (configs[1], [2], [4]): GPU path vs the CPU oracle on sampled queries of the
full-size synthetic corpora, plus size-independent properties on whole batches.  These are the same legs bench.py
prints as `parity` / `supplements.*.parity`; here they gate the test suite.  Needs one B200 (C5: ~110 GB of HBM)."""
import gc

import numpy as np
import pytest
import torch

import mse_b200  # noqa: F401
import bench
from mse_b200 import _native, synthetic
from oracle import sampled

pytestmark = pytest.mark.gpu


def _free():
    gc.collect()
    torch.cuda.empty_cache()


def test_c2_with_the_always_term_sampled_parity():
    """1M docs, vocab 200k, batch 1024 x (4 terms + the always-term in 95 % of the docs), top-1000."""
    r = bench.bm25_c2_supplement(torch.device("cuda", 0), 0, 6549.8, always=True, steps=3, warmup=1, parity_queries=16)
    assert r["parity"]["queries_failing"] == 0, r["parity"]
    assert r["parity"]["max_rel_err"] < 1e-5
    assert r["status_words"] == [0, 0, 0, 0]
    assert r["postings_looked_up_not_streamed_per_query"] > 5 * r["postings_streamed_per_query"]     # the always-term is looked up
    _free()


def test_c2_plain_sampled_parity():
    r = bench.bm25_c2_supplement(torch.device("cuda", 0), 0, 6549.8, always=False, steps=3, warmup=1, parity_queries=16)
    assert r["parity"]["queries_failing"] == 0, r["parity"]
    assert r["status_words"] == [0, 0, 0, 0] and r["postings_looked_up_not_streamed_per_query"] == 0
    _free()


def test_c3_dense_scan_10m_chunks_b1_and_b256_sampled_parity():
    """10M x 768 bf16 chunks, 2M docs, top-1000: the GEMV kernel (B=1) and the tcgen05 GEMM kernel (B=256) against the
    slab-wise float64-accumulate oracle over all 10M stored rows (D0 is a reconstruction: parity unpinned)."""
    r = bench.dense_c3_supplement(torch.device("cuda", 0), 0, 6549.8, 1602.5, steps=2, parity_queries=4)
    assert r["parity"]["queries_checked"] == 5 and r["parity"]["queries_failing"] == 0, r["parity"]
    assert r["parity"]["max_rel_err"] < 2e-3
    _free()


def test_c5_hybrid_10m_docs_sampled_parity_and_properties():
    """10M docs (always-term, ~1.9G postings) + 50M chunks: hybrid batch of 256 queries; 8 of them against the oracle
    pipeline, all of them for size-independent properties (descending fused scores in [0,1], unique docs, every result
    doc among the query's BM25 candidates, fused call == BM25 call + rerank call)."""
    dev = torch.device("cuda", 0)
    corpus = bench.build_corpus(dev, 0, 0, 1, bench.N_DOCS, bench.N_CHUNKS)
    nat = corpus.nat
    B, n_par = 256, 8
    host = bench.make_batches(corpus, 1, B, 4242)[0]
    par_ix, par_queries, _ = bench.cpu_hybrid_setup(corpus, dev, host, n_par)
    corpus.bm25 = None
    _free()
    bench.attach_dense(corpus, dev)
    q_off, q_term, q_tf, qv = host
    o_doc, o_score, o_orig, o_chunk, o_count, o_rows = nat.hybrid_search(q_off, q_term, q_tf, qv, 1000, 0.0, max_out=100)
    res = sampled.check_hybrid(par_ix, bench.fetch_rows_fn(corpus, dev), corpus.chunk_off_host, par_queries, qv[:n_par],
                               o_doc[:n_par], o_score[:n_par], o_count[:n_par], 1000, 100)
    assert res["queries_failing"] == 0, res
    b_doc, b_score, b_count = nat.bm25_search(q_off, q_term, q_tf, 1000, 0.0)
    st = nat.bm25_stats()
    assert st["rerun_queries"] == 0 and st["postings_looked_up"] > st["postings"]
    assert st["ranges"] == -(-bench.N_DOCS // 3072)                      # a corpus of this length gets the two-phase kernel (bm25_u16.cuh)
    nat.set_option("bm25_accum", 32)                                     # ... whose lists equal the fp32 kernel's, bit for bit
    f_doc, f_score, f_count = nat.bm25_search(q_off, q_term, q_tf, 1000, 0.0)
    assert nat.bm25_stats()["ranges"] == -(-bench.N_DOCS // 1536)
    assert np.array_equal(f_doc, b_doc) and np.array_equal(f_score, b_score) and np.array_equal(f_count, b_count)
    nat.set_option("bm25_accum", 0)
    bres = sampled.check_bm25(par_ix, par_queries, b_doc[:n_par], b_score[:n_par], b_count[:n_par], 1000)
    assert bres["queries_failing"] == 0, bres
    assert (o_count == 100).all() and (b_count == 1000).all()
    for i in range(B):
        s = o_score[i]
        assert (np.diff(s) <= 0).all() and s[0] <= 1.0 and s[-1] >= 0.0
        assert len(set(o_doc[i].tolist())) == 100 and set(o_doc[i].tolist()) <= set(b_doc[i].tolist())
        assert (np.diff(b_score[i]) <= 0).all() and b_score[i][-1] >= 0.0
    # stage-wise == fused, bit for bit
    cand_off = (np.arange(B + 1) * 1000).astype(np.int32)
    r2 = nat.rerank(cand_off, b_doc.reshape(-1), b_score.reshape(-1), qv, None, 0.15, 10, 100)
    assert np.array_equal(r2[0], o_doc) and np.array_equal(r2[1], o_score) and np.array_equal(r2[3], o_chunk)
    nat.close()
    corpus.emb = None
    _free()
