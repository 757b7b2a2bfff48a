"""The reference's other flat-similarity scans on the dense kernel.

* ``find_embedding_candidates`` -- ``GraphRetriever._find_embedding_candidates`` (reference
  ``graph/graph_retriever.py:153-170``): ``np.dot(embeddings, q)`` over all notes + ``argsort`` top-15.
* ``semantic_similarity_relations`` -- ``RelationExtractor.extract_semantic_similarity_relations`` (reference
  ``graph/relation_extractor.py:591-633``) with ``_compute_similarity_matrix`` (``:769-783``, N x N cosine with a
  zeroed diagonal) and ``_get_similarity_rank`` (``:785-792``): every pair i < j whose cosine reaches the
  threshold, with j's rank among i's neighbours.

Both run as batched exact top-k searches of ``engine.DenseIndex`` (fp32-exact mode: three bf16 planes on both sides),
so the N x N matrix is never materialised.  Floating point: similarities agree with numpy's fp32 matrix product to
1e-5 (the tolerance of the dense path); a pair or a rank can differ only where two similarities, or a similarity and
the threshold, are closer than that.
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence

import numpy as np

from .. import engine

_MAX_K = 256                      # ARB_MAX_TOPK


def find_embedding_candidates(embeddings, query_embedding, note_id_to_index: Dict[str, int], top_k: int = 15,
                              index: engine.DenseIndex = None) -> List[str]:
    """-> note ids of the ``top_k`` rows with the largest inner product, best first (:159-168).  ``index`` may hold the
    rows already (built with ``normalize=False``); otherwise a transient one is made from ``embeddings``."""
    if embeddings is None and index is None:
        return []                                                   # :155-156
    own = index is None
    if own:
        x = np.ascontiguousarray(embeddings, dtype=np.float32)
        index = engine.DenseIndex(x.shape[1], storage="fp32", normalize=False)
        index.add(x)
    try:
        n = index.ntotal
        k = min(int(top_k), n, _MAX_K)
        if k <= 0:
            return []
        q = np.ascontiguousarray(query_embedding, dtype=np.float32).reshape(1, -1)
        _, ids = index.search(q, k, query_planes=3)
    finally:
        if own:
            index.close()
    keys = list(note_id_to_index.keys())
    return [keys[i] for i in ids[0].tolist() if 0 <= i < len(keys)]   # :165-168


def semantic_similarity_relations(atomic_notes: Sequence[Dict[str, Any]], embeddings, similarity_threshold: float = 0.7,
                                  weight: float = 0.5) -> List[Dict[str, Any]]:
    """-> the relation dicts of extract_semantic_similarity_relations, in its order (i ascending, then j ascending)."""
    x = np.ascontiguousarray(embeddings, dtype=np.float32)
    n = len(atomic_notes)
    if x.shape[0] != n:                                             # :596-598
        return []
    if n < 2:
        return []
    index = engine.DenseIndex(x.shape[1], storage="fp32", normalize=True)   # :771-774 (zero rows stay zero)
    try:
        index.add(x)
        k = min(_MAX_K, n)
        scores, ids = index.search(x, k, query_planes=3)            # row i: its k nearest rows, itself among them
        # a row whose whole list is above the threshold may have more neighbours than the list holds: score it in full
        full = {}
        thr32 = np.float32(similarity_threshold)
        if k < n or similarity_threshold <= 0:
            # (threshold <= 0: the zeroed diagonal takes part in the ranking of the negative similarities, :780)
            crowded = np.arange(n) if similarity_threshold <= 0 else np.nonzero(scores[:, -1] >= thr32)[0]
            for lo in range(0, len(crowded), 64):
                rows = crowded[lo: lo + 64]
                cand = np.broadcast_to(np.arange(n, dtype=np.int64), (len(rows), n))
                sims = index.score_rows(x[rows], cand, query_planes=3, clamp_nonneg=False)
                for r, i in enumerate(rows.tolist()):
                    full[i] = sims[r]
    finally:
        index.close()
    relations = []
    for i in range(n):
        if i in full:
            s = full[i].copy()
            s[i] = 0.0                                              # :780 fill_diagonal(0)
            order = np.lexsort((-np.arange(n), -s))                 # descending similarity, larger index first on ties (:789)
            nb_ids, nb_s = order, s[order]
        else:
            keep = ids[i] != i
            nb_ids, nb_s = ids[i][keep], scores[i][keep]
            # the kernel breaks ties by ascending id; argsort()[::-1] (:789) puts the larger index first
            order = np.lexsort((-nb_ids, -nb_s))
            nb_ids, nb_s = nb_ids[order], nb_s[order]
        hit = np.nonzero((nb_s >= thr32) & (nb_ids > i))[0]
        for pos in sorted(hit.tolist(), key=lambda p: int(nb_ids[p])):      # j ascending (:605)
            j, sim = int(nb_ids[pos]), nb_s[pos]
            relations.append({
                "source_id": atomic_notes[i].get("note_id"), "target_id": atomic_notes[j].get("note_id"),
                "relation_type": "semantic_similarity", "weight": weight * float(sim),
                "metadata": {"cosine_similarity": float(sim), "similarity_rank": int(pos) + 1}})
    return relations
