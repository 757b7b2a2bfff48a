"""Candidate re-scoring of the live pipeline on the device.

Mirror of the fusion part of ``QueryProcessor._hybrid_search`` (reference
``query/query_processor.py:3681-3768``; defaults ``:392-402``): the caller supplies, per query, the candidate
notes with their vector similarities (``_calculate_vector_similarities`` :3543) and BM25 scores
(``bm25_scores`` :3707); the guards are evaluated on the host from the note text exactly like the reference
(:3717-3733, lower-cased substring tests), the arithmetic and the stable sort run in one CUDA kernel
(``csrc/fuse.cu`` ``rescore_kernel``) for a whole batch of queries.  Like the reference the candidates get a
``hybrid_score`` and come back sorted by it.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import _native as nat
from .._log import logger


def _guard_flags(candidates, must_have_terms, boost_entities, boost_predicates):
    miss, n_ent, n_pred = [], [], []
    for cand in candidates:
        content = cand.get("content", "").lower()
        miss.append(1 if must_have_terms and not any(t.lower() in content for t in must_have_terms) else 0)
        n_ent.append(sum(1 for e in (boost_entities or []) if e.lower() in content))
        n_pred.append(sum(1 for p in (boost_predicates or []) if p.lower() in content))
    return miss, n_ent, n_pred


class HybridRescorer:
    def __init__(self, fusion_method: str = "linear", vector_weight: Optional[float] = None,
                 bm25_weight: Optional[float] = None, rrf_k: float = 60, device: Optional[torch.device] = None):
        if fusion_method not in ("linear", "rrf"):
            raise ValueError(f"unknown fusion method {fusion_method!r}")
        self.fusion_method = fusion_method
        # query_processor.py:395-402
        self.vector_weight = vector_weight if vector_weight is not None else (0.7 if fusion_method == "linear" else 1.0)
        self.bm25_weight = bm25_weight if bm25_weight is not None else (0.3 if fusion_method == "linear" else 1.0)
        self.rrf_k = rrf_k
        self.device = device

    def rescore_batch(self, batch: Sequence[Dict[str, Any]]) -> List[List[Dict[str, Any]]]:
        """``batch``: one dict per query with ``candidates``, ``vector_scores``, ``bm25_scores`` and optionally
        ``must_have_terms`` / ``boost_entities`` / ``boost_predicates``.  One kernel launch for all of them."""
        nq = len(batch)
        if nq == 0:
            return []
        nc = max(len(b["candidates"]) for b in batch)
        if nc == 0:
            return [[] for _ in batch]
        vec = np.zeros((nq, nc), np.float64)
        bm = np.zeros((nq, nc), np.float64)
        flags = np.zeros((3, nq, nc), np.uint8)
        n_cand = np.zeros(nq, np.int32)
        for r, b in enumerate(batch):
            cands = b["candidates"]
            n = len(cands)
            n_cand[r] = n
            v = np.asarray(list(b["vector_scores"])[:n], dtype=np.float64)        # missing entries stay 0.0 (:3712-3713)
            s = np.asarray(list(b["bm25_scores"])[:n], dtype=np.float64)
            vec[r, :len(v)] = v
            bm[r, :len(s)] = s
            miss, n_ent, n_pred = _guard_flags(cands, b.get("must_have_terms"), b.get("boost_entities"), b.get("boost_predicates"))
            flags[0, r, :n], flags[1, r, :n], flags[2, r, :n] = miss, np.minimum(n_ent, 255), np.minimum(n_pred, 255)
        dev = self.device or torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):
            d_vec, d_bm = torch.from_numpy(vec).to(dev), torch.from_numpy(bm).to(dev)
            d_flags, d_n = torch.from_numpy(flags).to(dev), torch.from_numpy(n_cand).to(dev)
            out_s = torch.empty((nq, nc), dtype=torch.float64, device=dev)
            out_o = torch.empty((nq, nc), dtype=torch.int32, device=dev)
            nat.check(nat.lib.arb_rescore_candidates(
                nat.FUSE_LINEAR if self.fusion_method == "linear" else nat.FUSE_RRF, nq, nc, d_vec.data_ptr(), d_bm.data_ptr(),
                d_n.data_ptr(), d_flags[0].data_ptr(), d_flags[1].data_ptr(), d_flags[2].data_ptr(),
                float(self.vector_weight), float(self.bm25_weight), float(self.rrf_k), out_s.data_ptr(), out_o.data_ptr(),
                int(torch.cuda.current_stream(dev).cuda_stream)))
            scores, order = out_s.cpu().numpy(), out_o.cpu().numpy()
        results = []
        for r, b in enumerate(batch):
            cands = b["candidates"]
            for i, cand in enumerate(cands):
                cand["hybrid_score"] = float(scores[r, i])
            results.append([cands[i] for i in order[r, :len(cands)].tolist()])
        return results

    def hybrid_search(self, candidates: List[Dict[str, Any]], vector_scores: Sequence[float], bm25_scores_list: Sequence[float],
                      must_have_terms: List[str] = None, boost_entities: List[str] = None,
                      boost_predicates: List[str] = None) -> List[Dict[str, Any]]:
        """Single-query form with the reference's argument meaning (:3681-3684); never raises (:3766-3768):
        on failure the candidates come back ordered by vector similarity (_fallback_vector_search :3591)."""
        try:
            return self.rescore_batch([{"candidates": candidates, "vector_scores": vector_scores,
                                        "bm25_scores": bm25_scores_list, "must_have_terms": must_have_terms,
                                        "boost_entities": boost_entities, "boost_predicates": boost_predicates}])[0]
        except Exception as e:
            logger.error(f"Hybrid search failed: {e}")
            vs = list(vector_scores) + [0.0] * max(0, len(candidates) - len(vector_scores))
            return [c for c, _ in sorted(zip(candidates, vs), key=lambda x: x[1], reverse=True)]
