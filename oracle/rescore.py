"""Oracle: the live candidate re-scoring of the reference (TEST INFRASTRUCTURE ONLY).

Restates the fusion arithmetic of ``QueryProcessor._hybrid_search``
(``query/query_processor.py:3681-3768``): linear ``vector_weight * v + bm25_weight * b`` with the
retrieval guards (must-have terms: ``b *= 0.1``; every matching boost entity: ``v *= 1.2``; every
matching boost predicate: ``b *= 1.3``), or RRF over 0-based stable descending ranks with the
must-have guard applied to the fused score; then a stable descending sort.  PINNED: the method's own
source is executed by ``tests/golden/make_golden_rescore.py`` and the outputs are committed in
``tests/golden/rescore_golden.json``.  All arithmetic is Python float (fp64), which is what the
reference's numpy < 2 pin gives when float32 similarities meet Python floats.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def guard_flags(contents: Sequence[str], must_have_terms: Optional[Sequence[str]],
                boost_entities: Optional[Sequence[str]], boost_predicates: Optional[Sequence[str]]):
    """-> (must_miss, n_entities, n_predicates) per candidate; substring tests on lower-cased content (:3717-3733)."""
    miss, n_ent, n_pred = [], [], []
    for content in contents:
        c = (content or "").lower()
        miss.append(1 if must_have_terms and not any(t.lower() in c for t in must_have_terms) else 0)
        n_ent.append(sum(1 for e in (boost_entities or []) if e.lower() in c))
        n_pred.append(sum(1 for p in (boost_predicates or []) if p.lower() in c))
    return miss, n_ent, n_pred


def hybrid_search_rescore(vector: Sequence[float], bm25: Sequence[float], contents: Sequence[str], method: str,
                          vector_weight: float, bm25_weight: float, rrf_k: float = 60,
                          must_have_terms=None, boost_entities=None, boost_predicates=None) -> Tuple[List[int], List[float]]:
    """-> (candidate indices in result order, their hybrid_score)."""
    n = len(contents)
    vec = [vector[i] if i < len(vector) else 0.0 for i in range(n)]          # :3712-3713
    bm = [bm25[i] if i < len(bm25) else 0.0 for i in range(n)]
    miss, n_ent, n_pred = guard_flags(contents, must_have_terms, boost_entities, boost_predicates)
    score = [0.0] * n
    if method == "linear":                                                    # :3709-3737
        for i in range(n):
            v, b = vec[i], bm[i]
            if miss[i]:
                b *= 0.1
            for _ in range(n_ent[i]):
                v *= 1.2
            for _ in range(n_pred[i]):
                b *= 1.3
            score[i] = vector_weight * v + bm25_weight * b
    elif method == "rrf":                                                     # :3739-3757
        v_rank = {i: r for r, i in enumerate(sorted(range(n), key=lambda x: vec[x], reverse=True))}
        b_rank = {i: r for r, i in enumerate(sorted(range(n), key=lambda x: bm[x], reverse=True))}
        for i in range(n):
            s = vector_weight / (rrf_k + v_rank[i]) + bm25_weight / (rrf_k + b_rank[i])
            if miss[i]:
                s *= 0.1
            score[i] = s
    else:
        raise ValueError(method)
    order = sorted(range(n), key=lambda i: score[i], reverse=True)            # :3760 (stable)
    return order, [score[i] for i in order]
