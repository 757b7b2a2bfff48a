"""Mirror of the reference package ``utils`` restricted to ``bm25_search``."""
from . import bm25_search  # noqa: F401
