"""Exhaustive dense retriever — the ``Retriever`` the reference lists (``README.md:61``) and calls
(``search_api.py:60,87``: ``Retriever(embedder=, indexer=, db_path=)``, ``.quick_search(query, top_k=,
return_unique_docs=True)``) but no longer ships.  Reconstructed semantics:
inner product of the L2-normalised query (``indexer/embedder.py:54-61``) with every L2-normalised
chunk embedding (``indexer/indexer.py:165``), max-pooled per document, top-k unique documents,
ties to the lower doc id.  The scan runs on the GPU behind ``mse_dense_scan_batch``.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _native


class Retriever:
    def __init__(self, embedder: Optional[Callable[[str], np.ndarray]] = None, indexer=None, db_path: Optional[str] = None, *,
                 store=None, doc_ids: Optional[np.ndarray] = None, native: Optional[_native.NativeIndex] = None,
                 device: int = 0, dense_tables=None, doc_base: int = 0, chunk_base: int = 0, load: bool = True):
        if store is None and db_path is not None:
            from .store import open_store
            store = open_store(db_path, read_only=True)
        self.store = store
        self.embed = embedder
        self.doc_ids = np.asarray(doc_ids if doc_ids is not None else store.all_doc_ids(), dtype=np.int64)
        self.native = native if native is not None else _native.NativeIndex(device)
        self.doc_base = doc_base
        if load:
            dense = dense_tables if dense_tables is not None else store.load_dense(self.doc_ids)
            emb, off = dense.emb, dense.doc_chunk_off
            if not _native._is_torch(emb):
                emb = np.ascontiguousarray(emb, dtype=np.float32)
            if not _native._is_torch(off):
                off = np.ascontiguousarray(off, dtype=np.int64)
            self.native.dense_load(emb, off, doc_base=doc_base, chunk_base=chunk_base)

    def scan_batch(self, q_vecs, top_k: int = 1000, normalize: bool = True):
        """float32 [B,768] (numpy, or torch CUDA for the device-resident path) ->
        (doc index [B,k] global dense, score [B,k], count [B])."""
        if _native._is_torch(q_vecs):
            q = q_vecs.float()
            if normalize:
                q = q / q.norm(dim=1, keepdim=True)
            return self.native.dense_scan(q.contiguous(), top_k)
        q = np.asarray(q_vecs, dtype=np.float32)
        if normalize:
            q = q / np.linalg.norm(q, axis=1, keepdims=True)
        return self.native.dense_scan(np.ascontiguousarray(q, dtype=np.float32), top_k)

    def quick_search(self, query, top_k: int = 1000, return_unique_docs: bool = True) -> List[dict]:
        """One query (string through ``embedder``, or a vector) -> ``[{'doc_id', 'score'}]`` in rank order."""
        vec = self.embed(query) if isinstance(query, str) else np.asarray(query, dtype=np.float32)
        doc, score, count = self.scan_batch(vec[None, :], top_k=top_k)
        n = int(count[0])
        ids = self.doc_ids[doc[0, :n] - self.doc_base]
        return [{"doc_id": int(d), "score": float(s)} for d, s in zip(ids.tolist(), score[0, :n].tolist())]
