"""ano-rag_b200: B200-native (sm_100a) candidate scoring for AnoRAG.

Drop-in mirrors of the reference's hot-path API:

    ano_rag_b200.vector_store.VectorIndex / VectorRetriever   (vector_store/*.py)
    ano_rag_b200.utils.bm25_search                            (utils/bm25_search.py)
    ano_rag_b200.retrieval.hybrid_search.HybridSearcher       (retrieval/hybrid_search.py)

all backed by libanorag_b200.so (include/anorag_b200.h) through ctypes.  Importing
the package loads the shared object; there is no CPU implementation behind it.
"""
from . import _native  # noqa: F401  (fails loudly if the CUDA library is missing)
from . import engine  # noqa: F401

__all__ = ["engine"]
__version__ = "0.1.0"
