"""Oracle: the reference's flat-similarity scans outside the retriever (TEST INFRASTRUCTURE ONLY).

Restates ``GraphRetriever._find_embedding_candidates`` (``graph/graph_retriever.py:153-170``) and
``RelationExtractor.extract_semantic_similarity_relations`` with ``_compute_similarity_matrix`` /
``_get_similarity_rank`` (``graph/relation_extractor.py:591-633, 769-792``) in numpy.  PINNED: the methods' own source
is executed by ``tests/golden/make_golden_graph_sim.py``; outputs in ``tests/golden/graph_sim_golden.json``.
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence

import numpy as np


def find_embedding_candidates(embeddings, query_embedding, note_id_to_index: Dict[str, int], top_k: int = 15) -> List[str]:
    if embeddings is None:
        return []
    sims = np.dot(embeddings, query_embedding)                       # :159
    top = np.argsort(sims)[-top_k:][::-1]                            # :162
    keys = list(note_id_to_index.keys())
    return [keys[i] for i in top if i < len(keys)]                   # :165-168


def similarity_matrix(embeddings: np.ndarray) -> np.ndarray:
    norms = np.linalg.norm(embeddings, axis=1, keepdims=True)        # :771-774
    norms = np.where(norms == 0, 1, norms)
    x = embeddings / norms
    m = np.dot(x, x.T)                                               # :777
    np.fill_diagonal(m, 0)                                           # :780
    return m


def semantic_similarity_relations(atomic_notes: Sequence[Dict[str, Any]], embeddings: np.ndarray,
                                  similarity_threshold: float = 0.7, weight: float = 0.5) -> List[Dict[str, Any]]:
    if embeddings.shape[0] != len(atomic_notes):                     # :596-598
        return []
    m = similarity_matrix(embeddings)
    out = []
    for i in range(len(atomic_notes)):                               # :604-631
        order = None
        for j in range(i + 1, len(atomic_notes)):
            sim = m[i, j]
            if sim >= similarity_threshold:
                if order is None:
                    order = np.argsort(m[i])[::-1]                   # _get_similarity_rank :788-790
                out.append({"source_id": atomic_notes[i].get("note_id"), "target_id": atomic_notes[j].get("note_id"),
                            "relation_type": "semantic_similarity", "weight": weight * sim,
                            "metadata": {"cosine_similarity": float(sim),
                                         "similarity_rank": int(np.where(order == j)[0][0] + 1)}})
    return out
