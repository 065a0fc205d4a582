#!/usr/bin/env python
"""Sharded dense exhaustive scan (BASELINE.json configs[3]): 100M x 768 bf16 chunks (153.6 GB) sharded by
contiguous doc range over the GPUs of one box; every rank scans its shard, then an NCCL all-gather of the
per-rank top-k lists and a device-side merge.  Launch with torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/bench_dense_sharded.py --total-chunks 100000000 --batches 1,64

Prints one JSON line per batch size (rank 0): ms per batch = max over ranks (CUDA events bracketed by barriers).
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import mse_b200  # noqa
from mse_b200 import _native, synthetic
from mse_b200.sharding import ShardedSearcher

ap = argparse.ArgumentParser()
ap.add_argument("--total-chunks", type=int, default=100_000_000)
ap.add_argument("--chunks-per-doc", type=int, default=5)
ap.add_argument("--batches", default="1,64")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--top-k", type=int, default=1000)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
docs_total = a.total_chunks // a.chunks_per_doc
docs_local = docs_total // world
chunks_local = docs_local * a.chunks_per_doc
t0 = time.time()
d = synthetic.make_dense_corpus(docs_local, seed=1234 + rank, device=dev, dtype=torch.bfloat16, chunks_per_doc=a.chunks_per_doc)
nat = _native.NativeIndex(local)
nat.dense_load(d.emb, d.doc_chunk_off, doc_base=rank * docs_local, chunk_base=rank * chunks_local, borrow=True)
srch = ShardedSearcher(nat, rank, world)
peak = 6549.8
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
if rank == 0:
    print(f"# {world} ranks x {chunks_local} chunks ({chunks_local * 1536 / 1e9:.1f} GB each) ready in {time.time() - t0:.1f}s", file=sys.stderr)
for B in [int(x) for x in a.batches.split(",")]:
    q = torch.from_numpy(synthetic.make_query_vectors(B, seed=77, normalize=True)).to(dev)
    for _ in range(2):
        out = srch.dense_scan(q, a.top_k)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    nat.set_option("reset_timers", 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = srch.dense_scan(q, a.top_k)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    scan_ms, n = nat.kernel_time("dense_scan")
    scan_ms /= max(n, 1)
    t = torch.tensor([ms, scan_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, scan_ms = float(t[0]), float(t[1])
    # sanity: merged list is sorted and docs are unique
    doc, score, count = out
    ok = bool((score[:, :-1] >= score[:, 1:]).all().item()) and int(count.min().item()) == a.top_k
    if rank == 0:
        local_bytes = 2.0 * 768 * chunks_local
        print(json.dumps({"workload": f"sharded dense scan: {a.total_chunks} x 768 bf16 chunks over {world} GPU(s), B={B}, top-{a.top_k}",
                          "n_gpus": world, "batch": B, "ms_per_batch": ms, "queries_per_s": B / (ms / 1e3),
                          "scan_kernel_ms_max": scan_ms, "per_gpu_scan_GBps": local_bytes / (scan_ms * 1e-3) / 1e9,
                          "per_gpu_frac_of_hbm_peak": local_bytes / (scan_ms * 1e-3) / 1e9 / peak,
                          "exchange_and_merge_ms": ms - scan_ms, "sorted_and_full": ok}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
