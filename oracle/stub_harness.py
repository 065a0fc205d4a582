"""TEST INFRASTRUCTURE — hosts the UNMODIFIED reference modules on stand-in dependencies.

Only ``tests/`` and ``oracle/make_golden.py`` use this file, and only in the build
container where ``/root/reference`` is mounted.  It never runs on the GPU box and no
product code imports it.

What it does (recipe from SURVEY.md §8c): the reference's hot-path modules
(``indexer/bm25_indexer.py`` and ``reranker/reranker_api.py``) import ``duckdb``,
``spacy`` and ``sentence_transformers``, none of which exist here.  We register tiny
stand-ins in ``sys.modules`` and back the fake DuckDB connection with stdlib sqlite3:

* ``LOG`` is registered as ``float32(log10(x))`` — DuckDB's ``LOG`` is base 10 and the
  target column ``idf_score REAL`` is float32 (``bm25_indexer.py:110,140``).
* values written to ``bm25_corpus_stats.stat_value REAL`` are rounded to float32
  (``bm25_indexer.py:116-122,359-367``); sqlite ``REAL`` would keep float64.
* ``FIRST`` aggregate for ``reranker_api.py:39-41``.
* ``.df()`` builds a DataFrame, renames the repeated ``chunk_id`` column to
  ``chunk_id_1`` as DuckDB does, and decodes embedding BLOBs to float32[768].
* pandas >= 3 no longer hands the grouping column to ``groupby.apply``; the reference
  pins pandas 2.3.1 (``requirements.txt``), so a shim restores that behaviour.

Nothing of the reference is copied: the modules are loaded from where they lie.
"""
from __future__ import annotations

import asyncio
import importlib.util
import math
import os
import sqlite3
import sys
import types
from contextlib import contextmanager

import numpy as np

REFERENCE_ROOT = os.environ.get("MSE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "indexer", "bm25_indexer.py"))


# --------------------------------------------------------------------------- sqlite stand-in
class _First:
    """FIRST(x) aggregate: first value seen in the group."""

    def __init__(self):
        self.v, self.seen = None, False

    def step(self, x):
        if not self.seen:
            self.v, self.seen = x, True

    def finalize(self):
        return self.v


class _Result:
    def __init__(self, cursor):
        self._c = cursor

    def fetchall(self):
        return self._c.fetchall()

    def fetchone(self):
        return self._c.fetchone()

    def df(self):
        import pandas as pd

        names, seen = [], {}
        for d in self._c.description:
            n = d[0]
            if n in seen:
                seen[n] += 1
                n = f"{n}_{seen[n]}"
            else:
                seen[n] = 0
            names.append(n)
        rows = self._c.fetchall()
        cols = {n: [r[i] for r in rows] for i, n in enumerate(names)}
        for n in names:
            if n.startswith("embedding"):
                cols[n] = [np.frombuffer(v, dtype=np.float32).copy() for v in cols[n]]
        return pd.DataFrame(cols, columns=names)


class FakeDuckConnection:
    """One shared in-memory sqlite database presented with DuckDB's connection API."""

    def __init__(self, raw: sqlite3.Connection):
        self.raw = raw

    def execute(self, query, params=None):
        p = list(params) if params is not None else []
        if "INTO bm25_corpus_stats" in query and len(p) == 2:
            p[1] = float(np.float32(p[1]))  # REAL column is float32 in DuckDB
        return _Result(self.raw.execute(query, p))

    def executemany(self, query, seq):
        self.raw.executemany(query, list(seq))
        return self

    def commit(self):
        if self.raw.in_transaction:
            self.raw.commit()

    def create_function(self, *a, **k):  # embedder.py registers a UDF; unused here
        return None

    def close(self):
        pass


def _new_sqlite() -> sqlite3.Connection:
    raw = sqlite3.connect(":memory:", check_same_thread=False, isolation_level=None)
    raw.create_function("LOG", 1, lambda x: float(np.float32(math.log10(x))), deterministic=True)
    raw.create_aggregate("FIRST", 1, _First)
    return raw


# --------------------------------------------------------------------------- spaCy / ST stand-ins
class _Tok:
    __slots__ = ("lemma_",)
    is_stop = False
    is_punct = False
    is_alpha = True

    def __init__(self, s):
        self.lemma_ = s


def _fake_nlp(text):
    return [_Tok(t) for t in text.split()]


class _FakeEncoder:
    """SentenceTransformer stand-in: ``encode(text)`` returns the vector registered for it."""

    vectors: dict = {}

    def __init__(self, *a, **k):
        pass

    def encode(self, text, **k):
        return np.asarray(_FakeEncoder.vectors[text], dtype=np.float32)


def _pandas2_groupby_apply_shim():
    import pandas as pd

    if int(pd.__version__.split(".")[0]) < 3:
        return None
    from pandas.core.groupby.generic import DataFrameGroupBy

    orig = DataFrameGroupBy.apply

    def apply(self, func, *args, **kwargs):
        parts = [func(g.copy(), *args, **kwargs) for _, g in self]
        return pd.concat(parts)

    DataFrameGroupBy.apply = apply
    return (DataFrameGroupBy, orig)


@contextmanager
def hosted_reference():
    """Context manager yielding a ``Hosted`` object with ``.raw`` (sqlite), ``.BM25``, ``.load_reranker()``."""
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    raw = _new_sqlite()
    saved_modules = {k: sys.modules.get(k) for k in (
        "duckdb", "duckdb.typing", "spacy", "spacy.cli", "sentence_transformers", "config",
        "indexer", "indexer.bm25_indexer")}
    saved_path = list(sys.path)
    saved_cwd = os.getcwd()

    duck = types.ModuleType("duckdb")
    duck.connect = lambda path=None, read_only=False: FakeDuckConnection(raw)
    duck_typing = types.ModuleType("duckdb.typing")
    duck_typing.VARCHAR = "VARCHAR"
    duck.typing = duck_typing
    spacy = types.ModuleType("spacy")
    spacy.load = lambda name: _fake_nlp
    spacy_cli = types.ModuleType("spacy.cli")
    spacy_cli.download = lambda name: None
    spacy.cli = spacy_cli
    st = types.ModuleType("sentence_transformers")
    st.SentenceTransformer = _FakeEncoder
    sys.modules.update({"duckdb": duck, "duckdb.typing": duck_typing, "spacy": spacy,
                        "spacy.cli": spacy_cli, "sentence_transformers": st})
    for k in ("config", "indexer", "indexer.bm25_indexer"):
        sys.modules.pop(k, None)
    sys.path.insert(0, REFERENCE_ROOT)
    shim = _pandas2_groupby_apply_shim()
    os.chdir(REFERENCE_ROOT)  # reranker_api.py loads "reranker/config.yaml" relative to CWD

    class Hosted:
        pass

    h = Hosted()
    h.raw = raw
    h.conn = FakeDuckConnection(raw)
    h.encoder = _FakeEncoder

    def load_bm25_module():
        spec = importlib.util.spec_from_file_location(
            "ref_bm25_indexer", os.path.join(REFERENCE_ROOT, "indexer", "bm25_indexer.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    def load_reranker():
        spec = importlib.util.spec_from_file_location(
            "ref_reranker_api", os.path.join(REFERENCE_ROOT, "reranker", "reranker_api.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    def run_rerank(mod, doc_ids, similarities, query, query_vec):
        _FakeEncoder.vectors[query] = query_vec
        req = mod.RerankRequest(doc_ids=[str(d) for d in doc_ids],
                                similarities=None if similarities is None else [float(s) for s in similarities],
                                query=query)
        return asyncio.run(mod.rerank(req))

    h.bm25_module = load_bm25_module()
    h.BM25 = h.bm25_module.BM25
    h.load_reranker = load_reranker
    h.run_rerank = run_rerank
    try:
        yield h
    finally:
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
        if shim:
            shim[0].apply = shim[1]
        for k, v in saved_modules.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        raw.close()


# --------------------------------------------------------------------------- table helpers
URLS_DDL = """CREATE TABLE IF NOT EXISTS urlsDB (id BIGINT PRIMARY KEY, url TEXT UNIQUE, title TEXT,
 text TEXT, lastFetch DOUBLE, incoming TEXT, domainLinkingDepth TINYINT, linkingDepth TINYINT, tueEngScore DOUBLE)"""
CHUNKS_DDL = "CREATE TABLE IF NOT EXISTS chunks_optimized(chunk_id BIGINT PRIMARY KEY, doc_id BIGINT, chunk_text TEXT)"
EMB_DDL = "CREATE TABLE IF NOT EXISTS embeddings(chunk_id BIGINT PRIMARY KEY, embedding BLOB)"


def create_urls(raw, rows):
    """rows: iterable of (id, url, title, text) — DDL as crawler/databaseManagement.py:18-51."""
    raw.execute(URLS_DDL)
    raw.executemany("INSERT INTO urlsDB (id, url, title, text) VALUES (?,?,?,?)", list(rows))


def create_chunks(raw, chunk_ids, chunk_doc, emb):
    """DDL as indexer/embedder.py:31-52 (FLOAT[768] stored as a float32 BLOB in the stand-in)."""
    raw.execute(CHUNKS_DDL)
    raw.execute(EMB_DDL)
    raw.executemany("INSERT INTO chunks_optimized VALUES (?,?,?)",
                    [(int(c), int(d), f"chunk {int(c)}") for c, d in zip(chunk_ids, chunk_doc)])
    raw.executemany("INSERT INTO embeddings VALUES (?,?)",
                    [(int(c), np.asarray(e, dtype=np.float32).tobytes()) for c, e in zip(chunk_ids, emb)])


def bulk_load_bm25_tables(raw, doc_ids, doc_len, term_names, term_off, post_doc, post_tf):
    """Fill the four bm25_* tables directly (same result as build_index, without sqlite's
    32 766-parameter limit).  term_off/post_doc/post_tf are CSR by term; post_doc holds doc ids."""
    n = len(doc_ids)
    raw.executemany("INSERT OR REPLACE INTO bm25_doc_stats (doc_id, doc_length) VALUES (?,?)",
                    [(int(d), int(l)) for d, l in zip(doc_ids, doc_len)])
    tf_rows, ts_rows = [], []
    for t, name in enumerate(term_names):
        a, b = int(term_off[t]), int(term_off[t + 1])
        if b == a:
            continue
        for d, f in zip(post_doc[a:b], post_tf[a:b]):
            tf_rows.append((int(d), name, int(f)))
        ts_rows.append((name, b - a, int(np.sum(post_tf[a:b]))))
    raw.executemany("INSERT OR REPLACE INTO bm25_term_freq (doc_id, term, freq) VALUES (?,?,?)", tf_rows)
    raw.executemany("INSERT OR REPLACE INTO bm25_term_stats (term, doc_freq, total_freq) VALUES (?,?,?)", ts_rows)
    return n
