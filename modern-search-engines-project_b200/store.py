"""Index store adapters: read the reference's table layout into arrays.

The reference keeps everything in one DuckDB file (``config.py:3``): ``urlsDB``
(``crawler/databaseManagement.py:18-51``), the four ``bm25_*`` tables
(``indexer/bm25_indexer.py:82-128``) and ``chunks_optimized`` / ``embeddings``
(``indexer/embedder.py:31-52``).  ``SqlStore`` runs the load queries against any connection
that offers ``execute(sql, params).fetchall()`` — a real ``duckdb`` connection when that module is
importable, or ``sqlite3`` (the SQL used here is portable).  ``ArrayStore`` wraps arrays that
are already in memory (synthetic corpora).
"""
from __future__ import annotations

import os
import sqlite3
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class Bm25Tables:
    """CSR-by-term view of bm25_term_freq + the per-doc / per-term / corpus statistics."""
    terms: List[str]
    term_off: np.ndarray        # int64 [V+1]
    post_doc: np.ndarray        # int32 [P] dense doc index
    post_tf: np.ndarray         # int32 [P]
    doc_ids: np.ndarray         # int64 [N] ascending urlsDB ids
    doc_len: np.ndarray         # int32 [N]
    idf: np.ndarray             # float32 [V]  idf_score verbatim
    total_freq: np.ndarray      # int64 [V]
    avgdl: float
    total_docs: float


@dataclass
class DenseTables:
    emb: np.ndarray             # float32 (or torch bf16) [n_chunks, 768]
    chunk_ids: np.ndarray       # int64 [n_chunks] ascending, doc-contiguous
    doc_chunk_off: np.ndarray   # int64 [N+1] over the same dense doc index as Bm25Tables.doc_ids


class SqlStore:
    def __init__(self, conn, owns: bool = False):
        self.conn = conn
        self._owns = owns

    # -- helpers
    def _all(self, sql, params=()):
        return self.conn.execute(sql, list(params)).fetchall()

    def _columns(self, sql, params=()) -> List[np.ndarray]:
        """Result of a query with numeric columns as one numpy array per column.  A DuckDB cursor hands the columns
        over without a Python object per row (``fetchnumpy``); anything else (sqlite3) goes through ``fetchall``."""
        cur = self.conn.execute(sql, list(params))
        if hasattr(cur, "fetchnumpy"):
            d = cur.fetchnumpy()
            return [np.ma.getdata(v) if isinstance(v, np.ma.MaskedArray) else np.asarray(v) for v in d.values()]
        rows = cur.fetchall()
        n_cols = len(cur.description) if getattr(cur, "description", None) else (len(rows[0]) if rows else 0)
        if not rows:
            return [np.zeros(0, dtype=np.int64) for _ in range(n_cols)]
        return [np.asarray(c) for c in zip(*rows)]

    def has_table(self, name: str) -> bool:
        try:
            self.conn.execute(f"SELECT 1 FROM {name} LIMIT 1").fetchall()
            return True
        except Exception:
            return False

    # -- BM25 tables
    def bm25_fingerprint(self) -> str:
        """Cheap description of the state of the bm25_* tables: row counts, the newest ``processed_at`` /
        ``last_updated`` stamps (``bm25_indexer.py:90,105,115``) and the corpus statistics.  Two equal fingerprints
        mean the loaded CSR index (and an on-disk cache of it) is still current; `build_index` changes at least one
        of the fields (every run re-stamps ``bm25_corpus_stats`` and recomputes all idf values, ``:346-369``)."""
        d = self._all("SELECT COUNT(*), MAX(processed_at) FROM bm25_doc_stats")[0]
        t = self._all("SELECT COUNT(*), MAX(last_updated), SUM(doc_freq) FROM bm25_term_stats")[0]
        c = sorted((str(r[0]), repr(float(r[1])), str(r[2])) for r in
                   self._all("SELECT stat_name, stat_value, last_updated FROM bm25_corpus_stats"))
        return repr((int(d[0]), str(d[1]), int(t[0]), str(t[1]), int(t[2] or 0), c))

    def load_bm25(self, doc_ids: Optional[np.ndarray] = None) -> Bm25Tables:
        stats = dict(self._all("SELECT stat_name, stat_value FROM bm25_corpus_stats"))
        trows = self._all("SELECT term, doc_freq, total_freq, idf_score FROM bm25_term_stats ORDER BY term")
        drows = self._all("SELECT doc_id, doc_length FROM bm25_doc_stats ORDER BY doc_id")
        terms = [r[0] for r in trows]
        ids = np.asarray([r[0] for r in drows], dtype=np.int64)
        if doc_ids is not None:          # unified dense index (docs without tokens get length 0)
            all_ids = np.union1d(np.asarray(doc_ids, dtype=np.int64), ids)
        else:
            all_ids = ids
        doc_len = np.zeros(len(all_ids), dtype=np.int32)
        doc_len[np.searchsorted(all_ids, ids)] = np.asarray([r[1] for r in drows], dtype=np.int32)
        # postings as three integer columns: the term is replaced by its rank in `ORDER BY term` inside the database
        # (the same ordering the `terms` list above was read in), so no Python object is made per posting; postings of
        # a term without a bm25_term_stats row drop out of the join — such a term is unknown to search() anyway
        # (bm25_indexer.py:412-432).  The (term, doc) sort is done in numpy.
        cols = self._columns("SELECT s.rn, f.doc_id, f.freq FROM bm25_term_freq f JOIN "
                             "(SELECT term, ROW_NUMBER() OVER (ORDER BY term) - 1 AS rn FROM bm25_term_stats) s ON f.term = s.term")
        pt = cols[0].astype(np.int64)
        raw_doc = cols[1].astype(np.int64)
        pf = cols[2].astype(np.int32)
        # the reference's candidate query inner-joins bm25_doc_stats (bm25_indexer.py:442-444): a posting whose doc has
        # no doc_stats row never reaches the scoring loop.  Drop such rows here instead of attributing them to a
        # neighbouring doc (a partly built or inconsistent index).
        if len(ids):
            pos = np.minimum(np.searchsorted(ids, raw_doc), len(ids) - 1)
            ok = ids[pos] == raw_doc
        else:
            ok = np.zeros(len(raw_doc), dtype=bool)
        if not bool(np.all(ok)):
            pt, raw_doc, pf = pt[ok], raw_doc[ok], pf[ok]
        pd_ = np.searchsorted(all_ids, raw_doc).astype(np.int64)
        order = np.lexsort((pd_, pt))
        term_off = np.zeros(len(terms) + 1, dtype=np.int64)
        np.add.at(term_off, pt + 1, 1)
        term_off = np.cumsum(term_off)
        idf = np.asarray([(r[3] if r[3] is not None else 0.0) for r in trows], dtype=np.float32)
        return Bm25Tables(terms, term_off, pd_[order].astype(np.int32), pf[order], all_ids, doc_len, idf,
                          np.asarray([r[2] for r in trows], dtype=np.int64),
                          float(np.float32(stats.get("avg_doc_length", 1.0))), float(np.float32(stats.get("total_docs", 1))))

    # -- dense tables
    def dense_fingerprint(self) -> str:
        """Row counts and id range of the chunk / embedding tables (``indexer/embedder.py:31-52``; they carry no
        timestamps: chunks are only ever appended with increasing ``chunk_id``)."""
        c = self._all("SELECT COUNT(*), MAX(chunk_id) FROM chunks_optimized")[0]
        e = self._all("SELECT COUNT(*), MAX(chunk_id) FROM embeddings")[0]
        return repr((int(c[0]), int(c[1] or 0), int(e[0]), int(e[1] or 0)))

    def load_dense(self, doc_ids: np.ndarray) -> DenseTables:
        sql = ("SELECT c.doc_id, c.chunk_id, e.embedding FROM chunks_optimized c "
               "JOIN embeddings e ON c.chunk_id = e.chunk_id ORDER BY c.doc_id, c.chunk_id")
        doc_ids = np.asarray(doc_ids, dtype=np.int64)
        cur = self.conn.execute(sql, [])
        if hasattr(cur, "fetch_arrow_table"):             # DuckDB: FLOAT[768] arrives as one Arrow (fixed-size) list column
            return dense_from_arrow(cur.fetch_arrow_table(), doc_ids)
        rows = cur.fetchall()
        cd = np.asarray([r[0] for r in rows], dtype=np.int64)
        keep = np.isin(cd, doc_ids)
        chunk_ids = np.asarray([r[1] for r in rows], dtype=np.int64)[keep]
        emb = np.empty((int(keep.sum()), 768), dtype=np.float32)
        j = 0
        for r, k in zip(rows, keep):
            if k:
                v = r[2]
                emb[j] = np.frombuffer(v, dtype=np.float32) if isinstance(v, (bytes, memoryview)) else np.asarray(v, dtype=np.float32)
                j += 1
        counts = np.bincount(np.searchsorted(doc_ids, cd[keep]), minlength=len(doc_ids))
        off = np.zeros(len(doc_ids) + 1, dtype=np.int64)
        off[1:] = np.cumsum(counts)
        return DenseTables(emb, chunk_ids, off)

    # -- urlsDB
    def all_doc_ids(self) -> np.ndarray:
        return np.asarray([r[0] for r in self._all("SELECT id FROM urlsDB ORDER BY id")], dtype=np.int64)

    def urls(self, ids: Optional[Sequence[int]] = None) -> Dict[int, str]:
        if ids is None:
            return {int(r[0]): r[1] for r in self._all("SELECT id, url FROM urlsDB")}
        return {int(k): v[0] for k, v in self._fetch_by_ids("url", ids).items()}

    def _fetch_by_ids(self, cols: str, ids: Sequence[int], chunk: int = 900) -> Dict[int, tuple]:
        out: Dict[int, tuple] = {}
        ids = [int(i) for i in ids]
        for a in range(0, len(ids), chunk):
            part = ids[a:a + chunk]
            ph = ",".join("?" for _ in part)
            for r in self._all(f"SELECT id, {cols} FROM urlsDB WHERE id IN ({ph})", part):
                out[int(r[0])] = tuple(r[1:])
        return out

    def documents(self, ids: Sequence[int]) -> Dict[int, Tuple[str, str]]:
        """id -> (title, text) — the snippet query of bm25_indexer.py:494-501."""
        return self._fetch_by_ids("title, text", ids)

    def documents_full(self, ids: Sequence[int]) -> Dict[int, Tuple[str, str, str]]:
        """id -> (title, url, text) — what reranker_api.py:38-41 selects."""
        return self._fetch_by_ids("title, url, text", ids)

    def iter_documents(self):
        for r in self._all("SELECT id, title, text FROM urlsDB ORDER BY id"):
            yield int(r[0]), r[1], r[2]

    # -- writing the bm25_* tables (index build, bm25_indexer.py:82-128,283-367)
    def write_bm25(self, doc_rows, tf_rows, term_rows, avgdl: float, total_docs: int, idf_of):
        ex = self.conn.execute
        ex("CREATE TABLE IF NOT EXISTS bm25_doc_stats (doc_id INTEGER PRIMARY KEY, doc_length INTEGER, processed_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP)")
        ex("CREATE TABLE IF NOT EXISTS bm25_term_freq (doc_id INTEGER, term TEXT, freq INTEGER, PRIMARY KEY (doc_id, term))")
        ex("CREATE TABLE IF NOT EXISTS bm25_term_stats (term TEXT PRIMARY KEY, doc_freq INTEGER, total_freq INTEGER, idf_score REAL, last_updated TIMESTAMP DEFAULT CURRENT_TIMESTAMP)")
        ex("CREATE TABLE IF NOT EXISTS bm25_corpus_stats (stat_name TEXT PRIMARY KEY, stat_value REAL, last_updated TIMESTAMP DEFAULT CURRENT_TIMESTAMP)")
        self.conn.executemany("INSERT OR REPLACE INTO bm25_doc_stats (doc_id, doc_length) VALUES (?, ?)", list(doc_rows))
        self.conn.executemany("INSERT OR REPLACE INTO bm25_term_freq (doc_id, term, freq) VALUES (?, ?, ?)", list(tf_rows))
        self.conn.executemany("INSERT OR REPLACE INTO bm25_term_stats (term, doc_freq, total_freq, idf_score) VALUES (?, ?, ?, ?)",
                              [(t, df, tot, idf_of(df)) for t, df, tot in term_rows])
        for k, v in (("avg_doc_length", float(np.float32(avgdl))), ("total_docs", float(np.float32(total_docs)))):
            ex("INSERT OR REPLACE INTO bm25_corpus_stats (stat_name, stat_value) VALUES (?, ?)", [k, v])
        if hasattr(self.conn, "commit"):
            self.conn.commit()

    def close(self):
        if self._owns:
            try:
                self.conn.close()
            except Exception:
                pass


@dataclass
class ArrayStore:
    """In-memory store for synthetic corpora (integer-named terms, no text)."""
    bm25: Optional[Bm25Tables] = None
    dense: Optional[DenseTables] = None
    url_list: Optional[List[str]] = None        # parallel to bm25.doc_ids / dense doc index
    titles: Optional[List[str]] = None
    texts: Optional[List[str]] = None
    doc_id_array: Optional[np.ndarray] = None

    def _ids(self) -> np.ndarray:
        if self.doc_id_array is not None:
            return self.doc_id_array
        return self.bm25.doc_ids

    def has_table(self, name: str) -> bool:
        return {"bm25_term_freq": self.bm25 is not None, "chunks_optimized": self.dense is not None,
                "urlsDB": True}.get(name, False)

    def load_bm25(self, doc_ids=None) -> Bm25Tables:
        return self.bm25

    def load_dense(self, doc_ids) -> DenseTables:
        return self.dense

    def all_doc_ids(self) -> np.ndarray:
        return self._ids()

    def _index_of(self, ids):
        all_ids = self._ids()
        pos = np.searchsorted(all_ids, np.asarray(ids, dtype=np.int64))
        ok = (pos < len(all_ids)) & (all_ids[np.minimum(pos, len(all_ids) - 1)] == np.asarray(ids, dtype=np.int64))
        return pos, ok

    def urls(self, ids=None) -> Dict[int, str]:
        if self.url_list is None:
            return {}
        if ids is None:
            return {int(d): u for d, u in zip(self._ids().tolist(), self.url_list)}
        pos, ok = self._index_of(ids)
        return {int(d): self.url_list[p] for d, p, k in zip(ids, pos, ok) if k}

    def documents(self, ids) -> Dict[int, Tuple[str, str]]:
        pos, ok = self._index_of(ids)
        out = {}
        for d, p, k in zip(ids, pos, ok):
            if k:
                out[int(d)] = (self.titles[p] if self.titles else "", self.texts[p] if self.texts else "")
        return out

    def documents_full(self, ids) -> Dict[int, Tuple[str, str, str]]:
        pos, ok = self._index_of(ids)
        out = {}
        for d, p, k in zip(ids, pos, ok):
            if k:
                out[int(d)] = (self.titles[p] if self.titles else "", self.url_list[p] if self.url_list else "",
                               self.texts[p] if self.texts else "")
        return out

    def close(self):
        pass


# ---- on-disk cache of the loaded arrays ---------------------------------------------
def save_bm25_cache(path: str, t: Bm25Tables, fingerprint: str = "") -> None:
    """Persist the CSR arrays beside the database so the next start-up skips the SQL scan.  ``fingerprint``
    (`SqlStore.bm25_fingerprint`) records the table state the arrays were read from: `load_bm25_cache` returns None
    when the tables have changed since (or on a size mismatch)."""
    np.savez(path, term_off=t.term_off, post_doc=t.post_doc, post_tf=t.post_tf, doc_ids=t.doc_ids, doc_len=t.doc_len,
             idf=t.idf, total_freq=t.total_freq, scalars=np.asarray([t.avgdl, t.total_docs], dtype=np.float64),
             terms=np.asarray(t.terms if t.terms is not None else [], dtype=str),     # fixed-width unicode: no pickle
             fingerprint=np.asarray(fingerprint))


def load_bm25_cache(path: str, expect_docs: Optional[int] = None, expect_terms: Optional[int] = None,
                    fingerprint: Optional[str] = None) -> Optional[Bm25Tables]:
    if not os.path.exists(path):
        return None
    try:
        z = np.load(path, allow_pickle=False)          # the cache sits beside the database: never unpickle it
        if fingerprint is not None and ("fingerprint" not in z.files or str(z["fingerprint"]) != fingerprint):
            return None
        terms = [str(x) for x in z["terms"]]
    except ValueError:                                 # a cache written with pickled object arrays: treat as stale
        return None
    t = Bm25Tables(terms or None, z["term_off"], z["post_doc"], z["post_tf"], z["doc_ids"], z["doc_len"], z["idf"],
                   z["total_freq"], float(z["scalars"][0]), float(z["scalars"][1]))
    if expect_docs is not None and len(t.doc_ids) != expect_docs:
        return None
    if expect_terms is not None and len(t.term_off) - 1 != expect_terms:
        return None
    return t


def dense_from_arrow(table, doc_ids: np.ndarray, dim: int = 768) -> DenseTables:
    """(doc_id, chunk_id, embedding) Arrow table ordered by (doc_id, chunk_id) -> DenseTables without a Python object
    per row: the list column's child values are one contiguous float buffer."""
    import pyarrow as pa
    doc_ids = np.asarray(doc_ids, dtype=np.int64)
    cd = table.column(0).to_numpy().astype(np.int64)
    chunk_ids = table.column(1).to_numpy().astype(np.int64)
    col = table.column(2).combine_chunks()
    if isinstance(col, pa.ChunkedArray):
        col = col.chunk(0) if col.num_chunks else pa.array([], type=pa.list_(pa.float32()))
    n = len(col)
    if not pa.types.is_fixed_size_list(col.type):
        lengths = np.diff(col.offsets.to_numpy())
        if n and not np.all(lengths == dim):
            raise ValueError(f"embedding rows must have {dim} values (found lengths {np.unique(lengths)[:4]})")
    flat = col.flatten().to_numpy(zero_copy_only=False).astype(np.float32, copy=False)
    emb = flat.reshape(n, dim) if n else np.zeros((0, dim), dtype=np.float32)
    keep = np.isin(cd, doc_ids)
    counts = np.bincount(np.searchsorted(doc_ids, cd[keep]), minlength=len(doc_ids))
    off = np.zeros(len(doc_ids) + 1, dtype=np.int64)
    off[1:] = np.cumsum(counts)
    return DenseTables(np.ascontiguousarray(emb[keep]), chunk_ids[keep], off)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 bit patterns (uint16), round to nearest even — the conversion `mse_dense_load` applies to
    float32 input on the device (`__float2bfloat16_rn`), so a cache written here uploads to the same table."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = (u + (((u >> 16) & 1) + 0x7FFF)) >> 16
    r[nan] = (u[nan] >> 16) | 0x0040                      # quiet NaN, payload kept
    return r.astype(np.uint16)


def save_dense_cache(path: str, d: DenseTables, fingerprint: str = "") -> None:
    """Persist the chunk table as bf16 (half the size of the FLOAT[768] column) with its chunk ids and doc offsets."""
    emb = d.emb
    if hasattr(emb, "detach"):                            # torch tensor (bf16 or fp32)
        import torch
        bits = emb.detach().to(torch.bfloat16).cpu().contiguous().view(torch.int16).numpy().view(np.uint16)
    else:
        bits = f32_to_bf16_bits(np.asarray(emb, dtype=np.float32))
    np.savez(path, emb_bf16=bits, chunk_ids=np.asarray(d.chunk_ids, dtype=np.int64),
             doc_chunk_off=np.asarray(d.doc_chunk_off, dtype=np.int64), fingerprint=np.asarray(fingerprint))


def load_dense_cache(path: str, expect_docs: Optional[int] = None, fingerprint: Optional[str] = None) -> Optional[DenseTables]:
    """-> DenseTables whose ``emb`` is a torch bf16 tensor (uploaded as is), or None when absent / stale."""
    if not os.path.exists(path):
        return None
    z = np.load(path, allow_pickle=False)
    if fingerprint is not None and str(z["fingerprint"]) != fingerprint:
        return None
    off = z["doc_chunk_off"]
    if expect_docs is not None and len(off) - 1 != expect_docs:
        return None
    import torch
    emb = torch.from_numpy(z["emb_bf16"].view(np.int16)).view(torch.bfloat16)
    return DenseTables(emb, z["chunk_ids"], off)


def open_store(db_path: str, read_only: bool = True):
    """``duckdb.connect(db_path, read_only=...)`` as the reference does (bm25_indexer.py:69) when
    duckdb is importable; an sqlite3 file otherwise (same SQL)."""
    try:
        import duckdb  # type: ignore
        return SqlStore(duckdb.connect(db_path, read_only=read_only), owns=True)
    except ImportError:
        pass
    if not os.path.exists(db_path) and read_only:
        raise FileNotFoundError(db_path)
    with open(db_path, "rb") as f:
        head = f.read(16)
    if not head.startswith(b"SQLite format 3"):
        raise RuntimeError(f"{db_path} is not an sqlite file and the duckdb module is not installed")
    return SqlStore(sqlite3.connect(db_path, check_same_thread=False), owns=True)
