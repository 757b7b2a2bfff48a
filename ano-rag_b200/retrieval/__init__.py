"""Mirror of the reference package ``retrieval`` restricted to ``hybrid_search``."""
from .hybrid_search import HybridSearcher, create_hybrid_searcher  # noqa: F401
