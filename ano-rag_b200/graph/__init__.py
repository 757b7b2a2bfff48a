from .similarity import find_embedding_candidates, semantic_similarity_relations  # noqa: F401
