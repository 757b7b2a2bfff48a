"""Mirror of the reference package ``vector_store`` (hot-path classes only)."""
from .vector_index import VectorIndex
from .retriever import VectorRetriever

__all__ = ["VectorIndex", "VectorRetriever"]
