"""Oracle: dense/BM25/graph/path score fusion (TEST INFRASTRUCTURE ONLY).

Restates ``retrieval/hybrid_search.py:13-103`` (``HybridSearcher``).  PINNED: the
reference file loads by path; fixtures under tests/golden/ hold its outputs and
the tests compare this restatement with them (and with the live file when
``/root/reference`` exists).

One deliberate difference, documented in DESIGN.md: in ``linear`` mode the
reference iterates a Python ``set`` of ids (``hybrid_search.py:61,85``) so the
order of EXACTLY tied ``final_similarity`` values is hash-order dependent.  The
oracle (and the CUDA path) order exact ties by first appearance in the
dense -> bm25 -> graph -> path argument traversal; comparisons against the
reference are therefore made modulo permutations inside groups of exactly equal
``final_similarity``.  ``rrf`` mode has a deterministic order in the reference and
is reproduced exactly.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

SOURCES = ("dense", "bm25", "graph", "path")


class HybridSearcherOracle:
    def __init__(self, config: Dict[str, Any]):
        cfg = config if isinstance(config, dict) else config.load_config()   # :17
        r_cfg = cfg.get("retrieval", {})
        h_cfg = r_cfg.get("hybrid", {})
        self.candidate_pool = r_cfg.get("candidate_pool", 50)                # :20
        self.enabled = h_cfg.get("enabled", True)
        self.fusion_method = h_cfg.get("fusion_method", "linear")
        self.weights = h_cfg.get("weights", {})
        self.rrf_k = h_cfg.get("rrf_k", 60)

    @staticmethod
    def _normalize(scores: Dict[Any, float]) -> Dict[Any, float]:            # :26-32
        if not scores:
            return {}
        max_score = max(scores.values())
        if max_score == 0:
            return {k: 0.0 for k in scores}
        return {k: v / max_score for k, v in scores.items()}

    def fuse(self, dense=None, bm25=None, graph=None, path=None) -> List[Dict[str, Any]]:
        if not self.enabled:                                                 # :46-47
            return []
        lists = {"dense": dense or [], "bm25": bm25 or [], "graph": graph or [], "path": path or []}
        sources = {k: {nid: s for nid, s in v} for k, v in lists.items()}    # :54-59 last dup wins
        w = {k: self.weights.get(k, 0.0) for k in SOURCES}

        def entry(nid, final):
            return {
                "note_id": nid,
                "scores": {k: sources[k].get(nid) for k in SOURCES},
                "final_similarity": final,
                "tags": {"source": "graph" if nid in sources["graph"] else "semantic",
                         "is_bridge": nid in sources["path"]},
            }

        results = []
        if self.fusion_method == "rrf":                                      # :64-82
            ranks: Dict[Any, float] = {}
            for key in ("dense", "bm25", "graph"):
                sorted_ids = sorted(sources[key].items(), key=lambda x: x[1], reverse=True)
                for rank, (nid, _) in enumerate(sorted_ids, start=1):
                    ranks.setdefault(nid, 0.0)
                    ranks[nid] += w[key] / (self.rrf_k + rank)
            for nid, score in ranks.items():
                results.append(entry(nid, score + w["path"] * sources["path"].get(nid, 0.0)))
        else:                                                                # :83-100
            normed = {k: (self._normalize(v) if k != "path" else v) for k, v in sources.items()}
            order: Dict[Any, None] = {}
            for k in SOURCES:                     # first-appearance traversal (see module doc)
                for nid in sources[k]:
                    order.setdefault(nid, None)
            for nid in order:
                final = (w["dense"] * normed["dense"].get(nid, 0.0)
                         + w["bm25"] * normed["bm25"].get(nid, 0.0)
                         + w["graph"] * normed["graph"].get(nid, 0.0)
                         + w["path"] * normed["path"].get(nid, 0.0))
                results.append(entry(nid, final))
        results.sort(key=lambda x: x["final_similarity"], reverse=True)      # :102 (stable)
        return results[: self.candidate_pool]                                # :103


def canonical(results: List[Dict[str, Any]]):
    """Order-insensitive-within-exact-ties view of a fuse() result.

    Returns a list of (final_similarity, frozenset of (note_id, scores-tuple, tags-tuple))
    groups so two results can be compared modulo the reference's hash-order ties.
    NOTE: the candidate_pool cut can split a tie group; callers compare the pool
    boundary group by subset rather than equality.
    """
    groups = []
    for r in results:
        item = (r["note_id"], tuple(r["scores"][k] for k in SOURCES),
                (r["tags"]["source"], r["tags"]["is_bridge"]))
        if groups and groups[-1][0] == r["final_similarity"]:
            groups[-1][1].add(item)
        else:
            groups.append((r["final_similarity"], {item}))
    return groups


def same_modulo_ties(a: List[Dict[str, Any]], b: List[Dict[str, Any]], full_a=None, full_b=None) -> bool:
    """True if ``a`` and ``b`` hold bit-identical final scores position by position
    and identical entries up to permutation inside exact-tie groups.  The last
    group may be cut by candidate_pool: there each side's members only have to come
    from the same score level, which equal score sequences already guarantee when
    the uncut entry sets are unknown, so membership is checked as 'every entry of
    a's last group has b's last score and vice versa'."""
    if len(a) != len(b):
        return False
    if [r["final_similarity"] for r in a] != [r["final_similarity"] for r in b]:
        return False
    ga, gb = canonical(a), canonical(b)
    if len(ga) != len(gb):
        return False
    for i, ((sa, ia), (sb, ib)) in enumerate(zip(ga, gb)):
        if sa != sb:
            return False
        if i < len(ga) - 1 and ia != ib:
            return False
        if i == len(ga) - 1 and len(ia) != len(ib):
            return False
    return True
