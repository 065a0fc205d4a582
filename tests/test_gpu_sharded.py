"""Multi-GPU test of the sharded entry points (mse_comm_*, mse_bm25_search_sharded, mse_hybrid_search_sharded,
mse_dense_scan_sharded): one process per GPU, corpus sharded by doc range, NCCL inside the library.  Every rank
compares ITS block of the results with a single-GPU index over the whole corpus — they must be identical bit for bit
(a document's score is formed on one rank from global statistics with the same operations as on one GPU).
Skipped on a box with one GPU (run: gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu)."""
import os

import numpy as np
import pytest
import torch

import mse_b200  # noqa: F401

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n_docs, fail_path):
    try:
        import torch.distributed as dist
        from mse_b200 import _native, synthetic
        from mse_b200.reranker import url_groups
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
        box = [_native.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        c = synthetic.make_bm25_corpus(n_docs, vocab=5000, mean_len=64, seed=31, always_frac=0.95)
        counts = synthetic.make_chunk_counts(n_docs, 31)
        off = np.zeros(n_docs + 1, dtype=np.int64); off[1:] = np.cumsum(counts)
        emb = synthetic.dense_rows(0, int(off[-1]), 31, device="cpu", dtype=torch.float32).numpy()
        grp = url_groups(synthetic.make_urls(np.arange(1, n_docs + 1), dup_frac=0.05, seed=31))
        # single-GPU reference over the whole corpus (on this rank's GPU)
        full = _native.NativeIndex(rank)
        full.bm25_load(c.term_off.numpy(), c.post_doc.numpy(), c.post_tf.numpy(), c.doc_len.numpy(), c.idf.numpy(), c.avgdl)
        full.dense_load(emb, off)
        full.set_url_groups(grp)
        full.set_option("bm25_class_term", c.always_term)
        # this rank's shard
        lo, hi = rank * n_docs // world, (rank + 1) * n_docs // world
        pd, pt, toff = c.post_doc.numpy(), c.post_tf.numpy(), c.term_off.numpy()
        keep = (pd >= lo) & (pd < hi)
        term_of = np.repeat(np.arange(c.n_terms), np.diff(toff))
        s_off = np.zeros(c.n_terms + 1, dtype=np.int64); s_off[1:] = np.cumsum(np.bincount(term_of[keep], minlength=c.n_terms))
        sh = _native.NativeIndex(rank)
        sh.bm25_load(s_off, (pd[keep] - lo).astype(np.int32), pt[keep].copy(), c.doc_len.numpy()[lo:hi].copy(), c.idf.numpy(), c.avgdl, doc_base=lo)
        sh.dense_load(emb[off[lo]:off[hi]].copy(), (off[lo:hi + 1] - off[lo]).copy(), doc_base=lo, chunk_base=int(off[lo]))
        sh.set_url_groups(grp)                               # GLOBAL groups on every rank
        sh.set_option("bm25_class_term", c.always_term)
        sh.comm_init(box[0], rank, world)
        bq = 24
        GB = bq * world
        q_off, q_term, q_tf = synthetic.make_bm25_queries(c, GB, min_rank=8, seed=77, add_always=True)
        qv = synthetic.make_query_vectors(GB, seed=78)
        d = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (q_off, q_term, q_tf, qv))
        S = int(q_off[-1])
        status = torch.zeros(4, dtype=torch.int32, device=dev)
        blk = slice(rank * bq, (rank + 1) * bq)
        # ---- BM25: automatic shard list length, then a forced unsafe cut, then full lists
        ref = full.bm25_search(q_off, q_term, q_tf, 300, 0.0)
        for m in (0, 300):
            got = sh.bm25_search_sharded(d[0], d[1], d[2], S, 300, 0.0, shard_list_len=m, status=status)
            torch.cuda.synchronize()
            st = status.cpu().tolist()
            assert st[0] == 0 and st[1] == 0, st
            if st[2] == 0 or m == 300:
                assert st[2] == 0
                for name, g, r in zip(("doc", "score", "count"), got, ref):
                    assert np.array_equal(g.cpu().numpy(), r[blk]), (name, m)
        got = sh.bm25_search_sharded(d[0], d[1], d[2], S, 300, 0.0, shard_list_len=4, status=status)      # 4 entries per shard cannot cover a top-300
        torch.cuda.synchronize()
        flag = torch.tensor([int(status.cpu()[2].item())], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        assert flag.item() == 1
        # ---- hybrid
        href = full.hybrid_search(q_off, q_term, q_tf, qv, 1000, 0.0, max_out=100)
        hgot = sh.hybrid_search_sharded(*d, S, 1000, 0.0, shard_list_len=1000, max_out=100, status=status)
        torch.cuda.synchronize()
        assert status.cpu().tolist() == [0, 0, 0, 0]
        for name, g, r in zip(("doc", "score", "orig", "chunk", "count", "rows"), hgot, href):
            assert np.array_equal(g.cpu().numpy(), r[blk]), name
        hgot = sh.hybrid_search_sharded(*d, S, 1000, 0.0, max_out=100, status=status)                       # automatic list length
        torch.cuda.synchronize()
        st = status.cpu().tolist()
        assert st[0] == 0 and st[1] == 0
        if st[2] == 0:
            for name, g, r in zip(("doc", "score", "orig", "chunk", "count", "rows"), hgot, href):
                assert np.array_equal(g.cpu().numpy(), r[blk]), name
        # ---- dense scan (replicated result)
        qd = torch.from_numpy(synthetic.make_query_vectors(9, seed=5, normalize=True)).to(dev)
        sref = full.dense_scan(qd, 50)
        sgot = sh.dense_scan_sharded(qd, 50, status=status)
        torch.cuda.synchronize()
        for g, r in zip(sgot, sref):
            assert torch.equal(g, r)
        dist.barrier()
        sh.close(); full.close()
        dist.destroy_process_group()
    except BaseException as e:  # noqa: BLE001 - reported to the parent through a file (spawn swallows tracebacks of killed peers)
        import traceback
        with open(fail_path + f".{rank}", "w") as f:
            f.write(traceback.format_exc())
        raise


def test_sharded_calls_equal_single_gpu(tmp_path):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 500)
    fail = str(tmp_path / "fail")
    try:
        mp.spawn(_worker, args=(world, port, 30_000, fail), nprocs=world, join=True)
    except Exception:
        msgs = [open(fail + f".{r}").read() for r in range(world) if os.path.exists(fail + f".{r}")]
        raise AssertionError("\n".join(msgs) or "worker failed")
