"""world_size-2 gloo test of the multi-GPU exchange step on CPU tensors: all-gather (and the query-owner
all-to-all) of per-rank top-k lists + the merge rule == the unsharded oracle.  The per-rank scoring is
done by the oracle here (no GPU in this container); the collective plumbing is the code under test."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mse_b200  # noqa: F401
    import mse_testlib as helpers
    from mse_b200 import sharding, store
    from mse_b200.bm25_indexer import shard_bounds, slice_bm25_tables
    from oracle import bm25_oracle as bo
    ix, j, _ = helpers.load_bm25_small()
    t = store.Bm25Tables(ix.terms, ix.term_off, ix.post_doc, ix.post_tf, ix.doc_ids, ix.doc_len, ix.idf,
                         np.zeros(ix.n_terms, np.int64), ix.avgdl, ix.total_docs)
    b = shard_bounds(np.bincount(ix.post_doc, minlength=ix.n_docs), world)
    lo, hi = b[rank], b[rank + 1]
    s = slice_bm25_tables(t, lo, hi)
    local = bo.Bm25Arrays(s.term_off, s.post_doc, s.post_tf, s.doc_len, s.idf, s.avgdl, s.total_docs, s.doc_ids, ix.terms, ix.term_index)
    queries = [q["query"].split() for q in j["searches"]]
    k = 20
    doc = torch.full((len(queries), k), -1, dtype=torch.int32); score = torch.zeros((len(queries), k)); count = torch.zeros(len(queries), dtype=torch.int32)
    for i, q in enumerate(queries):
        # a term is valid when it exists GLOBALLY (replicated dictionary), even with no local posting
        terms = [w for w in q if w in ix.term_index]
        res = bo.search_fast(local, [ix.term_index[w] for w in terms], top_k=k, min_score=-100.0) if terms else []
        res = [(d, sc) for d, sc in res]
        # search_fast drops terms with no local postings, which is exactly "empty local segment"
        for r, (d, sc) in enumerate(res):
            doc[i, r] = d + lo; score[i, r] = sc
        count[i] = len(res)
    g_doc, g_score, g_count = sharding.all_gather_topk(doc, score, count)
    m_doc, m_score, m_count = sharding.merge_topk_host(g_doc.numpy(), g_score.numpy(), g_count.numpy(), k)
    # the truncated-first exchange must give the same answer (slack 0 forces cuts and exercises the exact fallback)
    host_merge = lambda a, b, c, kk: sharding.merge_topk_host(a.numpy(), b.numpy(), c.numpy(), kk)
    for slack, extra in ((2.0, 32), (1.2, 0), (0.3, 0)):     # full lists / mostly safe cuts / cuts that force the fallback
        x_doc, x_score, x_count = sharding.exchange_topk(doc, score, count, k, host_merge, slack=slack, extra=extra)
        assert np.array_equal(np.asarray(x_doc), m_doc) and np.array_equal(np.asarray(x_count), m_count)
        assert np.array_equal(np.asarray(x_score), m_score)
    # query-owner exchange (all-to-all): every rank receives the shard lists of ITS block of the batch; lists cut to
    # m entries stand for a shard-local search with top_k = m; an unsafe cut is repeated with full lists
    nq = (len(queries) // world) * world
    bq = nq // world
    blk = slice(rank * bq, (rank + 1) * bq)
    outcomes = []
    for m_cut in (k, 12, 3):
        cut = lambda t: t[:nq, :m_cut].contiguous()
        o_doc, o_score, o_count, unsafe = sharding.exchange_topk_owner(cut(doc), cut(score), torch.clamp(count[:nq], max=m_cut), k, host_merge)
        outcomes.append(bool(unsafe))
        if bool(unsafe):
            o_doc, o_score, o_count, again = sharding.exchange_topk_owner(doc[:nq].contiguous(), score[:nq].contiguous(), count[:nq].contiguous(), k, host_merge)
            assert not bool(again)
        assert np.array_equal(np.asarray(o_doc), m_doc[blk]) and np.array_equal(np.asarray(o_count), m_count[blk])
        assert np.array_equal(np.asarray(o_score), m_score[blk])
    assert outcomes[0] is False and outcomes[2] is True      # full lists are always safe; 2 x 3 entries cannot cover k = 20
    if rank == 0:
        ok = True
        for i, q in enumerate(queries):
            ref = bo.search_fast(ix, q, top_k=k, min_score=-100.0)
            n = int(m_count[i])
            ok &= n == len(ref)
            ok &= m_doc[i, :n].tolist() == [d for d, _ in ref]
            ok &= np.allclose(m_score[i, :n], [s_ for _, s_ in ref], rtol=1e-6)
        out.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_merge_equals_unsharded_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True
