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


# ---------------------------------------------------------------------------------------------------
# QueryProcessor._enhanced_hybrid_search_v2 (query/query_processor.py:1089-1143), without the optional
# ListT5 stage (:1146-1204, a neural re-ranker outside the path).  PINNED by
# tests/golden/make_golden_rescore_stages.py -> tests/golden/rescore_stages_golden.json ("v2").
# ---------------------------------------------------------------------------------------------------
def section_multiplier(cand: dict, enabled: bool, rule: str, fallback_to_lexical: bool) -> float:
    """_apply_section_filtering (:1235-1257)."""
    if not enabled:
        return 1.0
    role = cand.get("section_role")
    if role:
        if rule == "main_entity_related":
            low = role.lower()
            return 1.0 if ("main" in low or "primary" in low) else 0.8
        return 1.0
    return 1.0 if fallback_to_lexical else 0.9


def _full_text(cand: dict) -> str:
    return f"{cand.get('title', '').lower()} {cand.get('content', '').lower()} {cand.get('raw_span', '').lower()}"


def lexical_multiplier(cand: dict, must_have_terms, enabled: bool, miss_penalty: float) -> float:
    """_apply_lexical_fallback (:1259-1281)."""
    if not enabled or not must_have_terms:
        return 1.0
    text = _full_text(cand)
    hits = sum(1 for t in must_have_terms if t.lower() in text)
    if hits == 0:
        return miss_penalty
    return 0.8 if hits < len(must_have_terms) / 2 else 1.0


def satisfies_must(cand: dict, must_have_terms) -> bool:
    """_satisfies_must_have_terms (:1283-1298)."""
    if not must_have_terms:
        return True
    text = _full_text(cand)
    return any(t.lower() in text for t in must_have_terms)


def capped_boost(cand: dict, terms, step: float, cap: float) -> float:
    """_calculate_entity_boost (:1300-1312, 1.2 / 2.0) and _calculate_predicate_boost (:1314-1326, 1.15 / 1.8)."""
    if not terms:
        return 1.0
    content = cand.get("content", "").lower()
    f = 1.0
    for t in terms:
        if t.lower() in content:
            f *= step
    return min(f, cap)


def enhanced_hybrid_search_v2(dense: Sequence[float], sparse: Sequence[float], cands: Sequence[dict], settings: dict,
                              must_have_terms=None, boost_entities=None, boost_predicates=None):
    """-> (kept candidate indices in result order, final_base_score of every candidate in input order)."""
    n = len(cands)
    score = [0.0] * n
    for i, cand in enumerate(cands):
        d = dense[i] if i < len(dense) else 0.0                                # :1105-1106
        s = sparse[i] if i < len(sparse) else 0.0
        f = 1.0 * d + 0.6 * s                                                  # :1109
        if settings["section_filtering_enabled"]:                              # :1112-1114
            f *= section_multiplier(cand, True, settings["section_filter_rule"], settings["fallback_to_lexical"])
        if settings["lexical_fallback_enabled"] and must_have_terms:           # :1117-1119
            f *= lexical_multiplier(cand, must_have_terms, True, settings["miss_penalty"])
        if f < settings["noise_threshold"] and not satisfies_must(cand, must_have_terms):   # :1122-1123
            f = 0.0
        if boost_entities:                                                     # :1126-1132
            f *= capped_boost(cand, boost_entities, 1.2, 2.0)
        if boost_predicates:
            f *= capped_boost(cand, boost_predicates, 1.15, 1.8)
        score[i] = f
    kept = [i for i in range(n) if score[i] > 0]                               # :1139
    kept.sort(key=lambda i: score[i], reverse=True)                            # :1142 (stable)
    return kept, score


# ---------------------------------------------------------------------------------------------------
# QueryProcessor._rerank_khop_candidates (:1549-1635) with _rrf_score (:854-861) and _compute_cov_cons
# (:843-852).  PINNED by the same generator ("khop").
# ---------------------------------------------------------------------------------------------------
def cov_cons(note: dict, path_entities) -> Tuple[float, int]:
    ne = set(note.get("entities") or [])
    pe = set(e.lower() for e in (path_entities or []))
    if not pe:
        return 0.0, 0
    cov = len({e.lower() for e in ne} & pe) / max(1, len(pe))
    text = f"{note.get('title', '')} {note.get('content', '')}".lower()
    return float(cov), (1 if any(e in text for e in pe) else 0)


def _cfg_float(d: dict, key: str, default: float) -> float:
    try:
        return float(d.get(key, default))
    except (TypeError, ValueError):
        return default


def rerank_khop(dense: Sequence[float], bm25: Sequence[float], focused: Sequence[float], cands: Sequence[dict],
                path_entities=None, relation_focus: str = "", fusion: Optional[dict] = None, features: Optional[dict] = None):
    """-> (candidate indices in result order, final_score of every candidate in input order).
    ``focused[i]`` is the similarity of candidate i with the focus query (:1612); it is only used when that
    query is non-empty (:1607, :1611)."""
    fusion, features = fusion or {}, features or {}
    dense_w = _cfg_float(fusion, "dense_weight", 1.0)                           # :1556-1583
    bm25_w = _cfg_float(fusion, "bm25_weight", 0.6)
    focused_w2 = _cfg_float(fusion, "focused_weight_hop2", 0.30)
    hop_decay = _cfg_float(features, "hop_decay", 0.85)
    cov_w = _cfg_float(features, "cov_weight", 0.10)
    cons_w = _cfg_float(features, "cons_weight", 0.05)
    rrf_lambda = _cfg_float(fusion, "rrf_lambda", 0.2)
    n = len(cands)
    if n == 0:
        return [], []

    def ranks_from(scores):                                                    # :1590-1593 (later duplicates overwrite)
        order = sorted(range(len(scores)), key=lambda i: scores[i], reverse=True)
        return {cands[i].get("note_id"): r + 1 for r, i in enumerate(order)}

    prf = [i for i in range(n) if cands[i].get("retrieval_method") == "prf_bridge"]
    prf.sort(key=lambda i: cands[i].get("similarity_score", 0.0), reverse=True)
    ranks = [ranks_from(list(dense)), ranks_from(list(bm25)),
             {cands[i].get("note_id"): r + 1 for r, i in enumerate(prf)}]       # :1595-1602
    rrf = {}
    for table in ranks:                                                        # _rrf_score, k = 50 (:854-861, :1603)
        for nid, r in table.items():
            rrf[nid] = rrf.get(nid, 0.0) + 1.0 / (50 + r)
    q_focus = " ".join([*list(set([e for e in (path_entities or [])][-2:])), relation_focus]).strip()   # :1605-1606
    score = [0.0] * n
    for i, cand in enumerate(cands):
        base = dense_w * (dense[i] if i < len(dense) else 0.0) + bm25_w * (bm25[i] if i < len(bm25) else 0.0)
        foc = focused[i] if q_focus else 0.0
        cov, cons = cov_cons(cand, path_entities)
        bonus = rrf.get(cand.get("note_id"), 0.0)
        hop = int(cand.get("hop_no", 1))
        fw = focused_w2 if hop == 2 else max(0.0, focused_w2 - 0.10 * (hop - 2))
        final = base + fw * foc + cov_w * cov + cons_w * cons + rrf_lambda * bonus       # :1625
        if hop > 1 and 0.0 < hop_decay < 1.0:
            final *= hop_decay ** (hop - 1)
        score[i] = float(final)
    order = sorted(range(n), key=lambda i: score[i], reverse=True)             # :1632 (stable)
    return order, score


def cluster_suppression(candidates, vecs, cos_threshold: float = 0.90, keep_per_cluster: int = 2):
    """``QueryProcessor._apply_cluster_suppression`` (query/query_processor.py:1637-1697) restated.

    Greedy near-duplicate suppression over a small candidate list: walk the candidates in order; an unused candidate i
    opens a cluster of itself and every later unused j with ``cos(v_i, v_j) >= cos_threshold`` (vectors divided by
    ``norm + 1e-8``, :1676-1682); the cluster is sorted by ``final_score`` (stable, descending, missing = 0.0) and its
    first ``keep_per_cluster`` members are kept, ALL members are marked used (:1684-1689); the kept list is finally
    sorted by ``final_score`` again (stable, :1691).  PINNED: tests/golden/cluster_suppression_golden.json.
    """
    import numpy as np
    kept, used = [], set()
    unit = [np.asarray(v) / (np.linalg.norm(v) + 1e-8) for v in vecs]
    for i, ci in enumerate(candidates):
        if i in used:
            continue
        cluster = [(i, ci)]
        for j in range(i + 1, len(candidates)):
            if j in used:
                continue
            if float(unit[i] @ unit[j]) >= cos_threshold:
                cluster.append((j, candidates[j]))
        cluster.sort(key=lambda x: x[1].get("final_score", 0.0), reverse=True)
        for k, c in cluster[:keep_per_cluster]:
            kept.append(c)
            used.add(k)
        for k, _ in cluster[keep_per_cluster:]:
            used.add(k)
    kept.sort(key=lambda x: x.get("final_score", 0.0), reverse=True)
    return kept
