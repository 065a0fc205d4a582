"""B200-native retrieval hot path (BM25 scorer, dense scan / rerank, fusion, top-k) behind the
query API of StephenTaf/Modern-Search-Engines-Project.  CUDA kernels live in ``csrc/`` behind the C
ABI of ``include/mse_b200.h``; this package is the host-side mirror of the reference interface."""
from . import _native, bm25_indexer, pipeline, reranker, retriever, sharding, store, synthetic  # noqa: F401
from ._native import NativeError, NativeIndex  # noqa: F401
from .bm25_indexer import BM25, bm25_from_arrays  # noqa: F401
from .reranker import Reranker, RerankRequest, RerankResponse  # noqa: F401
from .retriever import Retriever  # noqa: F401
