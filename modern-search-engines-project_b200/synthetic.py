"""Seeded synthetic corpora in the reference's table shapes.

The generators are written with torch ops so the same code runs on the host (small test
corpora) and on the GPU (the 1M-doc / 10M-chunk bench corpora, where a host generator would
take minutes).  They produce *index arrays*, i.e. what the loader would read out of
``bm25_term_freq`` / ``bm25_doc_stats`` / ``bm25_term_stats`` / ``bm25_corpus_stats``
(``/root/reference/indexer/bm25_indexer.py:82-128``) and ``chunks_optimized`` / ``embeddings``
(``/root/reference/indexer/embedder.py:31-52``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch


@dataclass
class SyntheticBm25:
    term_off: torch.Tensor      # int64 [V+1]
    post_doc: torch.Tensor      # int32 [P] dense doc index, ascending inside a term
    post_tf: torch.Tensor       # int32 [P]
    doc_len: torch.Tensor       # int32 [N]
    idf: torch.Tensor           # float32 [V]
    avgdl: float                # float32 value
    total_docs: float           # float32 value
    doc_ids: torch.Tensor       # int64 [N] ascending (1-based like the crawler)
    always_term: int = -1       # index of the "tübingen"-like term, -1 if absent

    @property
    def n_docs(self) -> int:
        return int(self.doc_len.numel())

    @property
    def n_terms(self) -> int:
        return int(self.term_off.numel() - 1)

    @property
    def n_postings(self) -> int:
        return int(self.post_doc.numel())


def zipf_cdf(vocab: int, s: float = 1.0, device="cpu") -> torch.Tensor:
    r = torch.arange(1, vocab + 1, dtype=torch.float64, device=device)
    p = r.pow(-s)
    return torch.cumsum(p / p.sum(), 0)


def idf_float32(total_docs: float, df: torch.Tensor) -> torch.Tensor:
    """float32(log10((N - df + 0.5)/(df + 0.5))) — bm25_indexer.py:138-141 (DuckDB LOG is base 10)."""
    d = df.to(torch.float64)
    return torch.log10((float(total_docs) - d + 0.5) / (d + 0.5)).to(torch.float32)


def make_bm25_corpus(n_docs: int, vocab: int = 200_000, mean_len: float = 256.0, sigma: float = 0.5,
                     zipf_s: float = 1.0, seed: int = 1234, device="cpu",
                     always_frac: float = 0.0, sort_limit: int = (1 << 31) - 1) -> SyntheticBm25:
    """Docs with log-normal lengths and iid Zipf tokens; ``always_frac`` > 0 adds one extra term
    (index ``vocab``) present in that fraction of docs with tf ~ 1 + Poisson(3) — the analogue of
    the "tübingen" term ``search_api.py:160-165`` appends to every query."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    mu = math.log(mean_len) - sigma * sigma / 2
    z = torch.randn(n_docs, generator=g, device=dev, dtype=torch.float64)
    L = torch.clamp(torch.round(torch.exp(mu + sigma * z)), min=8).to(torch.int64)
    cdf = zipf_cdf(vocab, zipf_s, dev)
    n_tok = int(L.sum().item())
    step = 1 << 26

    def slab_postings(d0: int, d1: int):
        """(term int64, doc int32, tf int32) of docs [d0, d1), sorted by (term, doc)."""
        Ls = L[d0:d1]
        nt = int(Ls.sum().item())
        doc_of = torch.repeat_interleave(torch.arange(d0, d1, device=dev, dtype=torch.int64), Ls)
        keys = torch.empty(nt, dtype=torch.int64, device=dev)
        for a in range(0, nt, step):             # slabs keep the fp64 uniforms small
            e = min(nt, a + step)
            u = torch.rand(e - a, generator=g, device=dev, dtype=torch.float64)
            t = torch.searchsorted(cdf, u).clamp_(max=vocab - 1)
            keys[a:e] = t * n_docs + doc_of[a:e]
        del doc_of
        keys = torch.sort(keys).values            # (the permutation is not needed: drop it right away)
        uniq, counts = torch.unique_consecutive(keys, return_counts=True)
        del keys
        return uniq // n_docs, (uniq % n_docs).to(torch.int32), counts.to(torch.int32)

    if n_tok < sort_limit:
        term, post_doc, post_tf = slab_postings(0, n_docs)
    else:
        # torch.sort takes at most INT_MAX elements: generate doc-range slabs of <= sort_limit/2 tokens, each sorted by
        # (term, doc), and interleave them term-major (slab order == doc order inside every term)
        csum = torch.cumsum(L, 0)
        bounds, lo_tok = [0], 0
        while bounds[-1] < n_docs:
            nxt = int(torch.searchsorted(csum, torch.tensor([lo_tok + max(1, sort_limit // 2)], device=dev)).item())
            nxt = max(min(nxt, n_docs), bounds[-1] + 1)
            bounds.append(nxt)
            lo_tok = int(csum[nxt - 1].item())
        slabs, dfs = [], []
        for d0, d1 in zip(bounds[:-1], bounds[1:]):
            t_s, d_s, f_s = slab_postings(d0, d1)
            dfs.append(torch.bincount(t_s, minlength=vocab))
            slabs.append((d_s, f_s))
            del t_s
        df_all = torch.stack(dfs).sum(0)
        t_off = torch.zeros(vocab + 1, dtype=torch.int64, device=dev)
        t_off[1:] = torch.cumsum(df_all, 0)
        n_post = int(t_off[-1].item())
        post_doc = torch.empty(n_post, dtype=torch.int32, device=dev)
        post_tf = torch.empty(n_post, dtype=torch.int32, device=dev)
        base = t_off[:-1].clone()
        ar = torch.arange(vocab, device=dev, dtype=torch.int64)
        for (d_s, f_s), df_s in zip(slabs, dfs):
            s_off = torch.cumsum(df_s, 0) - df_s
            t_s = torch.repeat_interleave(ar, df_s)
            dest = torch.arange(d_s.numel(), device=dev, dtype=torch.int64) - s_off[t_s] + base[t_s]
            post_doc[dest] = d_s
            post_tf[dest] = f_s
            base += df_s
            del t_s, dest
        del slabs
        term = torch.repeat_interleave(ar, df_all)
    n_terms = vocab
    doc_len = L.clone()
    always_term = -1
    if always_frac > 0:
        sel = torch.rand(n_docs, generator=g, device=dev) < always_frac
        docs = torch.nonzero(sel).flatten()
        lam = torch.full((docs.numel(),), 3.0, device=dev)
        tf = (1 + torch.poisson(lam, generator=g)).to(torch.int32)
        term = torch.cat([term, torch.full((docs.numel(),), vocab, dtype=torch.int64, device=dev)])
        post_doc = torch.cat([post_doc, docs.to(torch.int32)])
        post_tf = torch.cat([post_tf, tf])
        doc_len[docs] += tf.to(torch.int64)
        always_term = vocab
        n_terms = vocab + 1
    df = torch.bincount(term, minlength=n_terms)
    term_off = torch.zeros(n_terms + 1, dtype=torch.int64, device=dev)
    term_off[1:] = torch.cumsum(df, 0)
    avgdl = float(np.float32(doc_len.to(torch.float64).mean().item()))
    total = float(np.float32(n_docs))
    idf = idf_float32(total, df)
    return SyntheticBm25(term_off, post_doc, post_tf, doc_len.to(torch.int32), idf, avgdl, total,
                         torch.arange(1, n_docs + 1, dtype=torch.int64, device=dev), always_term)


def make_bm25_queries(corpus: SyntheticBm25, n_queries: int, terms_per_query: int = 4, min_rank: int = 64,
                      zipf_s: float = 1.0, seed: int = 1235, repeat_frac: float = 0.05,
                      add_always: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """CSR query batch (q_off int32[B+1], q_term int32, q_tf int32): ``terms_per_query`` distinct
    ranks drawn from the Zipf restricted to ranks >= ``min_rank``; ``repeat_frac`` of the queries
    use one term twice (qtf=2); terms that occur in no document are dropped (bm25_indexer.py:430)."""
    rng = np.random.Generator(np.random.Philox(seed))
    vocab = corpus.n_terms - (1 if corpus.always_term >= 0 else 0)
    lo = min(min_rank, max(0, vocab - terms_per_query))
    r = np.arange(lo + 1, vocab + 1, dtype=np.float64)
    p = r ** (-zipf_s)
    cdf = np.cumsum(p / p.sum())
    df = np.diff(corpus.term_off.cpu().numpy())
    q_off, q_term, q_tf = [0], [], []
    for _ in range(n_queries):
        seen = []
        while len(seen) < terms_per_query:
            t = lo + int(np.searchsorted(cdf, rng.random()))
            t = min(t, vocab - 1)
            if t not in seen:
                seen.append(t)
        tfs = [1] * len(seen)
        if rng.random() < repeat_frac:
            tfs[int(rng.integers(len(seen)))] = 2
        if add_always and corpus.always_term >= 0:
            seen.append(corpus.always_term); tfs.append(1)
        for t, f in zip(seen, tfs):
            if df[t] > 0:
                q_term.append(t); q_tf.append(f)
        q_off.append(len(q_term))
    return (np.asarray(q_off, dtype=np.int32), np.asarray(q_term, dtype=np.int32),
            np.asarray(q_tf, dtype=np.int32))


@dataclass
class SyntheticDense:
    emb: torch.Tensor            # bf16 or float32 [n_chunks, dim], L2-normalised rows
    doc_chunk_off: torch.Tensor  # int64 [N+1]
    chunk_ids: torch.Tensor      # int64 [n_chunks] (sequential from 0, doc-contiguous; indexer.py:108-111)


def make_chunk_counts(n_docs: int, seed: int = 1234, p: float = 0.2, cap: int = 32,
                      total: Optional[int] = None) -> np.ndarray:
    """Chunks per doc ~ min(Geometric(p), cap) (mean ~5; ~11 % exceed the rerank cap of 10).
    With ``total`` the tail is trimmed/padded so the counts sum to it exactly."""
    rng = np.random.Generator(np.random.Philox(seed + 7))
    n = np.minimum(rng.geometric(p, size=n_docs), cap).astype(np.int64)
    if total is not None:
        diff = int(total - n.sum())
        i = n_docs - 1
        while diff != 0:
            if diff > 0:
                n[i] += 1; diff -= 1
            elif n[i] > 1:
                n[i] -= 1; diff += 1
            i = i - 1 if i > 0 else n_docs - 1
    return n


def make_dense_corpus(n_docs: int, dim: int = 768, seed: int = 1234, device="cpu", dtype=torch.bfloat16,
                      chunks_per_doc: Optional[int] = None, total_chunks: Optional[int] = None) -> SyntheticDense:
    """Unit-norm Gaussian chunk embeddings, doc-contiguous.  ``chunks_per_doc`` fixes the count
    (C3: 5 chunks/doc); otherwise clipped-geometric counts."""
    dev = torch.device(device)
    if chunks_per_doc is not None:
        counts = np.full(n_docs, chunks_per_doc, dtype=np.int64)
    else:
        counts = make_chunk_counts(n_docs, seed, total=total_chunks)
    off = np.zeros(n_docs + 1, dtype=np.int64)
    off[1:] = np.cumsum(counts)
    n_chunks = int(off[-1])
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 11)
    emb = torch.empty((n_chunks, dim), dtype=dtype, device=dev)
    slab = 1 << 18
    for a in range(0, n_chunks, slab):
        e = min(n_chunks, a + slab)
        x = torch.randn((e - a, dim), generator=g, device=dev, dtype=torch.float32)
        x = x / x.norm(dim=1, keepdim=True)
        emb[a:e] = x.to(dtype)
    return SyntheticDense(emb, torch.from_numpy(off).to(dev), torch.arange(n_chunks, dtype=torch.int64, device=dev))


DENSE_SLAB_ROWS = 1 << 18


def dense_rows(row_lo: int, row_hi: int, seed: int = 1234, dim: int = 768, device="cpu", dtype=torch.bfloat16) -> torch.Tensor:
    """Rows [row_lo, row_hi) of the shardable synthetic chunk table: unit-norm Gaussian rows generated in slabs of
    ``DENSE_SLAB_ROWS`` rows, slab s from its own generator seeded with (seed, s) — so ANY rank (or the CPU oracle)
    regenerates exactly the rows it needs, whatever the sharding.  Values are identical on CPU
    and CUDA only up to the generator implementation: regenerate on the device type the table was built on."""
    dev = torch.device(device)
    out = torch.empty((max(0, row_hi - row_lo), dim), dtype=dtype, device=dev)
    s0, s1 = row_lo // DENSE_SLAB_ROWS, (max(row_hi, row_lo + 1) - 1) // DENSE_SLAB_ROWS
    for sl in range(s0, s1 + 1):
        a, e = sl * DENSE_SLAB_ROWS, (sl + 1) * DENSE_SLAB_ROWS
        lo, hi = max(a, row_lo), min(e, row_hi)
        if hi <= lo:
            continue
        g = torch.Generator(device=dev)
        g.manual_seed((seed + 11) * 1_000_003 + sl)
        x = torch.randn((DENSE_SLAB_ROWS, dim), generator=g, device=dev, dtype=torch.float32)[lo - a:hi - a]
        out[lo - row_lo:hi - row_lo] = (x / x.norm(dim=1, keepdim=True)).to(dtype)
    return out


def make_dense_shard(doc_chunk_off_global: np.ndarray, doc_lo: int, doc_hi: int, seed: int = 1234, device="cpu",
                     dtype=torch.bfloat16) -> SyntheticDense:
    """The chunk rows of docs [doc_lo, doc_hi) of the shardable table (``dense_rows``): local offsets, global chunk ids."""
    off = np.asarray(doc_chunk_off_global, dtype=np.int64)
    r0, r1 = int(off[doc_lo]), int(off[doc_hi])
    emb = dense_rows(r0, r1, seed, device=device, dtype=dtype)
    return SyntheticDense(emb, torch.from_numpy(off[doc_lo:doc_hi + 1] - r0).to(device),
                          torch.arange(r0, r1, dtype=torch.int64, device=device))


def make_query_vectors(n: int, dim: int = 768, seed: int = 1235, normalize: bool = False) -> np.ndarray:
    """Standard-normal query vectors, float32; un-normalised for the rerank path
    (reranker_api.py:355), normalised for the exhaustive scan (embedder.py:58)."""
    rng = np.random.Generator(np.random.Philox(seed + 3))
    q = rng.standard_normal((n, dim)).astype(np.float32)
    if normalize:
        q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q


def make_urls(doc_ids: np.ndarray, n_domains: int = 997, dup_frac: float = 0.02, seed: int = 1234):
    """``https://d{doc_id % n_domains}.example/{doc_id}``; ``dup_frac`` of the docs are ``?q=x``
    variants of the preceding doc's URL (exercises the URL-dedupe of reranker_api.py:38-47)."""
    rng = np.random.Generator(np.random.Philox(seed + 5))
    urls = []
    for i, d in enumerate(np.asarray(doc_ids).tolist()):
        if i > 0 and rng.random() < dup_frac:
            urls.append(urls[i - 1].split("?")[0] + f"?q={d}")
        else:
            urls.append(f"https://d{d % n_domains}.example/{d}")
    return urls
