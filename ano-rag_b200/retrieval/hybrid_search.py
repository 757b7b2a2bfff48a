"""HybridSearcher -- drop-in for reference retrieval/hybrid_search.py:13-115.

``fuse`` keeps the reference signature and result schema; the arithmetic (per-source
max-normalisation + weighted sum, or weighted RRF), the stable descending sort and
the candidate_pool cut run in one CUDA kernel (csrc/fuse.cu) in fp64, in the
reference's operation order.  The host only maps arbitrary hashable note ids to
integers and back (dict semantics of :54-59: the LAST score of a duplicated id, at
its FIRST position).  ``fuse_batch`` is the batched tensor path the pipeline uses.

Tie order: the reference's linear mode iterates a Python set (hash order); here
exactly tied final scores keep first-appearance order of the dense, bm25, graph,
path traversal.  rrf order is the reference's.
"""
from __future__ import annotations

import math
from typing import Any, Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np

from .. import engine
from .._log import logger

_SRC = ("dense", "bm25", "graph", "path")
MAX_PER_SOURCE = 256        # kernel limit (include/anorag_b200.h arb_fuse)


class HybridSearcher:
    """Fuse scores from multiple retrievers according to configuration."""

    def __init__(self, config: Dict[str, Any]):
        cfg = config if isinstance(config, dict) else config.load_config()      # :17
        r_cfg = cfg.get("retrieval", {})
        h_cfg = r_cfg.get("hybrid", {})
        self.candidate_pool = r_cfg.get("candidate_pool", 50)
        self.enabled = h_cfg.get("enabled", True)
        self.fusion_method = h_cfg.get("fusion_method", "linear")
        self.weights = h_cfg.get("weights", {})
        self.rrf_k = h_cfg.get("rrf_k", 60)

    # ------------------------------------------------------------------ reference-shaped API
    def fuse(self, dense: Optional[List[Tuple[Hashable, float]]] = None,
             bm25: Optional[List[Tuple[Hashable, float]]] = None,
             graph: Optional[List[Tuple[Hashable, float]]] = None,
             path: Optional[List[Tuple[Hashable, float]]] = None) -> List[Dict[str, Any]]:
        if not self.enabled:                                                     # :46-47
            return []
        lists = [dense or [], bm25 or [], graph or [], path or []]
        # dict construction of :54-59 -- first position, last score
        per_src: List[Dict[Hashable, float]] = [{nid: s for nid, s in lst} for lst in lists]
        if not any(per_src):
            return []
        for i, d in enumerate(per_src):
            if len(d) > MAX_PER_SOURCE:
                # the fusion kernel holds MAX_PER_SOURCE entries per source; the reference has no limit and never raises.
                # Keep the source's best MAX_PER_SOURCE by score (stable, dict order otherwise unchanged): their ranks and
                # the normalising maximum are what they were; only contributions of rank > MAX_PER_SOURCE are lost.
                logger.warning(f"fuse(): {_SRC[i]} list has {len(d)} ids; keeping its best {MAX_PER_SOURCE}")
                keep = set(sorted(d, key=lambda n: d[n], reverse=True)[:MAX_PER_SOURCE])
                per_src[i] = {n: v for n, v in d.items() if n in keep}
        code: Dict[Hashable, int] = {}
        for d in per_src:
            for nid in d:
                code.setdefault(nid, len(code))
        names = list(code)
        srcs = []
        for d in per_src:
            if not d:
                srcs.append(None)
                continue
            ids = np.fromiter((code[n] for n in d), dtype=np.int64, count=len(d))[None, :]
            sc = np.fromiter((float(v) for v in d.values()), dtype=np.float64, count=len(d))[None, :]
            srcs.append((ids, sc))
        pool = min(max(int(self.candidate_pool), 0), len(code))
        if pool == 0:
            return []
        o_ids, o_fin, _o_src, o_cnt = engine.fuse_batch(self.fusion_method, self.weights, self.rrf_k, pool,
                                                        dense=srcs[0], bm25=srcs[1], graph=srcs[2], path=srcs[3])
        results = []
        for j in range(int(o_cnt[0])):
            nid = names[int(o_ids[0, j])]
            results.append({
                "note_id": nid,
                "scores": {k: per_src[i].get(nid) for i, k in enumerate(_SRC)},   # raw values incl. ints, None if absent
                "final_similarity": float(o_fin[0, j]),
                "tags": {"source": "graph" if nid in per_src[2] else "semantic", "is_bridge": nid in per_src[3]},
            })
        return results

    # ------------------------------------------------------------------ batched path
    def fuse_tensors(self, dense=None, bm25=None, graph=None, path=None, pool: Optional[int] = None):
        """Each source: (ids int64 [nq,k], scores [nq,k]) numpy or CUDA tensors, id < 0 = empty slot.
        -> (ids [nq,pool], final fp64 [nq,pool], src fp64 [nq,pool,4] (NaN = absent), count [nq])."""
        pool = int(self.candidate_pool if pool is None else pool)
        return engine.fuse_batch(self.fusion_method, self.weights, self.rrf_k, pool, dense=dense, bm25=bm25,
                                 graph=graph, path=path)

    def fuse_batch(self, dense=None, bm25=None, graph=None, path=None) -> List[List[Dict[str, Any]]]:
        """Batched ``fuse`` over integer note ids; returns the reference's dict schema per query."""
        if not self.enabled:
            n = next(len(s[0]) for s in (dense, bm25, graph, path) if s is not None)
            return [[] for _ in range(n)]
        o_ids, o_fin, o_src, o_cnt = self.fuse_tensors(dense, bm25, graph, path)
        if hasattr(o_ids, "cpu"):
            o_ids, o_fin, o_src, o_cnt = (t.cpu().numpy() for t in (o_ids, o_fin, o_src, o_cnt))
        out = []
        ids_l, fin_l, src_l = o_ids.tolist(), o_fin.tolist(), o_src.tolist()
        for q in range(len(ids_l)):
            res = []
            for j in range(int(o_cnt[q])):
                sc = src_l[q][j]
                res.append({
                    "note_id": ids_l[q][j],
                    "scores": {k: (None if math.isnan(sc[i]) else sc[i]) for i, k in enumerate(_SRC)},
                    "final_similarity": fin_l[q][j],
                    "tags": {"source": "semantic" if math.isnan(sc[2]) else "graph", "is_bridge": not math.isnan(sc[3])},
                })
            out.append(res)
        return out


def create_hybrid_searcher(config: Dict[str, Any]) -> HybridSearcher:
    """Factory (reference :106-115)."""
    return HybridSearcher(config)
