"""The other two live fusion stages of the reference's QueryProcessor on the device.

* ``EnhancedHybridSearchV2`` -- ``QueryProcessor._enhanced_hybrid_search_v2`` (reference
  ``query/query_processor.py:1089-1143``; settings ``:559-568``): ``1.0 * dense + 0.6 * sparse``, section /
  lexical multipliers, noise-threshold cut, entity / predicate boosts, drop non-positive, stable sort.  The
  optional ListT5 stage that may follow (``:1146-1204``) is a neural re-ranker and stays with the caller.
* ``KhopReranker`` -- ``QueryProcessor._rerank_khop_candidates`` (``:1549-1635``): dense/BM25 base score plus
  focused similarity, entity coverage / consistency (``_compute_cov_cons`` ``:843-852``), RRF over the dense,
  BM25 and PRF ranks (``_rrf_score`` ``:854-861``, k = 50) and hop decay.

As in ``hybrid_rescore.py`` everything the reference reads from note TEXT is evaluated on the host with the
same lower-cased substring tests and handed to the kernel as per-candidate numbers; the arithmetic, the ranks
and the stable sorts of a whole batch of queries run in one launch (``csrc/fuse.cu`` ``rescore_stage_kernel``
behind ``arb_rescore_v2`` / ``arb_rescore_khop``).  The similarities are inputs: ``VectorRetriever.
calculate_vector_similarities`` and ``bm25_scores`` produce them for many queries at once.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import _native as nat
from .._log import logger


def _stream(dev) -> int:
    return int(torch.cuda.current_stream(dev).cuda_stream)


def _padded(rows: Sequence[Sequence[float]], lens: Sequence[int], nc: int, dtype=np.float64) -> np.ndarray:
    """Ragged per-query lists -> [nq, nc]; entries a list does not have stay 0 (the reference's `if i < len(...)`)."""
    out = np.zeros((len(rows), nc), dtype)
    for r, (row, n) in enumerate(zip(rows, lens)):
        row = list(row)[:n]
        out[r, :len(row)] = row
    return out


class EnhancedHybridSearchV2:
    def __init__(self, section_filtering_enabled: bool = True, section_filter_rule: str = "main_entity_related",
                 fallback_to_lexical: bool = True, lexical_fallback_enabled: bool = True, miss_penalty: float = 0.6,
                 noise_threshold: float = 0.20, device: Optional[torch.device] = None):
        # defaults: query_processor.py:559-568
        self.section_filtering_enabled = section_filtering_enabled
        self.section_filter_rule = section_filter_rule
        self.fallback_to_lexical = fallback_to_lexical
        self.lexical_fallback_enabled = lexical_fallback_enabled
        self.miss_penalty = miss_penalty
        self.noise_threshold = noise_threshold
        self.device = device

    # -- host-side text rules ------------------------------------------------------------------------
    def _section(self, cand: Dict[str, Any]) -> float:
        """:1235-1257"""
        role = cand.get("section_role")
        if role:
            if self.section_filter_rule != "main_entity_related":
                return 1.0
            role = role.lower()
            return 1.0 if "main" in role or "primary" in role else 0.8
        return 1.0 if self.fallback_to_lexical else 0.9

    @staticmethod
    def _searchable(cand: Dict[str, Any]) -> str:
        return " ".join(cand.get(f, "").lower() for f in ("title", "content", "raw_span"))

    def _lexical(self, text: str, terms: Sequence[str]) -> float:
        """:1259-1281"""
        hits = sum(1 for t in terms if t in text)
        if hits == 0:
            return self.miss_penalty
        return 0.8 if hits < len(terms) / 2 else 1.0

    @staticmethod
    def _boost(content: str, terms: Sequence[str], step: float, cap: float) -> float:
        """:1300-1326"""
        f = 1.0
        for t in terms:
            if t in content:
                f *= step
        return min(f, cap)

    # -- batched ----------------------------------------------------------------------------------
    def search_batch(self, batch: Sequence[Dict[str, Any]]) -> List[List[Dict[str, Any]]]:
        """``batch``: one dict per query with ``candidates``, ``dense_scores``, ``sparse_scores`` and optionally
        ``must_have_terms`` / ``boost_entities`` / ``boost_predicates``.  One launch; every candidate gets
        ``final_base_score`` / ``dense_score`` / ``sparse_score`` (:1134-1136), the kept ones come back sorted."""
        nq = len(batch)
        lens = [len(b["candidates"]) for b in batch]
        nc = max(lens, default=0)
        if nc == 0:
            return [[] for _ in batch]
        dense = _padded([b["dense_scores"] for b in batch], lens, nc)
        sparse = _padded([b["sparse_scores"] for b in batch], lens, nc)
        mult = np.ones((4, nq, nc), np.float64)                      # section, lexical, entity, predicate
        sat = np.ones((nq, nc), np.uint8)
        for r, b in enumerate(batch):
            must = [t.lower() for t in (b.get("must_have_terms") or [])]
            ents = [t.lower() for t in (b.get("boost_entities") or [])]
            preds = [t.lower() for t in (b.get("boost_predicates") or [])]
            for i, cand in enumerate(b["candidates"]):
                if self.section_filtering_enabled:
                    mult[0, r, i] = self._section(cand)
                if must:
                    text = self._searchable(cand)
                    if self.lexical_fallback_enabled:
                        mult[1, r, i] = self._lexical(text, must)
                    sat[r, i] = 1 if any(t in text for t in must) else 0
                if ents or preds:
                    content = cand.get("content", "").lower()
                    if ents:
                        mult[2, r, i] = self._boost(content, ents, 1.2, 2.0)
                    if preds:
                        mult[3, r, i] = self._boost(content, preds, 1.15, 1.8)
        dev = self.device or torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):
            d_dense, d_sparse = torch.from_numpy(dense).to(dev), torch.from_numpy(sparse).to(dev)
            d_mult, d_sat = torch.from_numpy(mult).to(dev), torch.from_numpy(sat).to(dev)
            d_n = torch.tensor(lens, dtype=torch.int32, device=dev)
            out_s = torch.empty((nq, nc), dtype=torch.float64, device=dev)
            out_o = torch.empty((nq, nc), dtype=torch.int32, device=dev)
            out_k = torch.empty((nq,), dtype=torch.int32, device=dev)
            nat.check(nat.lib.arb_rescore_v2(nq, nc, d_dense.data_ptr(), d_sparse.data_ptr(), d_n.data_ptr(),
                                             d_mult[0].data_ptr(), d_mult[1].data_ptr(), d_sat.data_ptr(),
                                             d_mult[2].data_ptr(), d_mult[3].data_ptr(), 1.0, 0.6, float(self.noise_threshold),
                                             out_s.data_ptr(), out_o.data_ptr(), out_k.data_ptr(), _stream(dev)))
            score, order, kept = out_s.cpu().numpy(), out_o.cpu().numpy(), out_k.cpu().numpy()
        results = []
        for r, b in enumerate(batch):
            cands = b["candidates"]
            for i, cand in enumerate(cands):
                cand["final_base_score"] = float(score[r, i])
                cand["dense_score"] = float(dense[r, i])
                cand["sparse_score"] = float(sparse[r, i])
            results.append([cands[i] for i in order[r, :kept[r]].tolist()])
        return results

    def enhanced_hybrid_search_v2(self, candidates: List[Dict[str, Any]], dense_scores: Sequence[float],
                                  sparse_scores: Sequence[float], must_have_terms: List[str] = None,
                                  boost_entities: List[str] = None, boost_predicates: List[str] = None) -> List[Dict[str, Any]]:
        """Single-query form (:1089-1093).  Like the reference it never raises: on failure the error is logged and
        the candidates come back as they were (:1206-1209)."""
        if not candidates:
            return []
        try:
            return self.search_batch([{"candidates": candidates, "dense_scores": dense_scores, "sparse_scores": sparse_scores,
                                       "must_have_terms": must_have_terms, "boost_entities": boost_entities,
                                       "boost_predicates": boost_predicates}])[0]
        except Exception as e:
            logger.error(f"Enhanced hybrid search v2 failed: {e}")
            return candidates


def _cfg_float(d: Dict[str, Any], key: str, default: float) -> float:
    try:
        return float(d.get(key, default))
    except (TypeError, ValueError):
        return default


class KhopReranker:
    RRF_K = 50                                                       # _rrf_score default, :854, :1603

    def __init__(self, fusion: Optional[Dict[str, Any]] = None, features: Optional[Dict[str, Any]] = None,
                 device: Optional[torch.device] = None):
        fusion, features = fusion or {}, features or {}              # hybrid_search.fusion / .features, :1555-1583
        self.dense_weight = _cfg_float(fusion, "dense_weight", 1.0)
        self.bm25_weight = _cfg_float(fusion, "bm25_weight", 0.6)
        self.focused_weight_hop2 = _cfg_float(fusion, "focused_weight_hop2", 0.30)
        self.rrf_lambda = _cfg_float(fusion, "rrf_lambda", 0.2)
        self.hop_decay = _cfg_float(features, "hop_decay", 0.85)
        self.cov_weight = _cfg_float(features, "cov_weight", 0.10)
        self.cons_weight = _cfg_float(features, "cons_weight", 0.05)
        self.device = device

    @staticmethod
    def focus_query(relation_focus: str, path_entities: Optional[Sequence[str]] = None) -> str:
        """The text whose similarity with each candidate is the ``focused`` input (:1605-1606): the last two path
        entities (deduplicated) and the relation phrase of the query."""
        return " ".join([*list(set(list(path_entities or [])[-2:])), relation_focus]).strip()

    @staticmethod
    def _cov_cons(note: Dict[str, Any], wanted: set):
        """:843-852"""
        if not wanted:
            return 0.0, 0
        have = {e.lower() for e in (note.get("entities") or [])}
        text = f"{note.get('title', '')} {note.get('content', '')}".lower()
        return len(have & wanted) / max(1, len(wanted)), (1 if any(e in text for e in wanted) else 0)

    def rerank_batch(self, batch: Sequence[Dict[str, Any]]) -> List[List[Dict[str, Any]]]:
        """``batch``: one dict per query with ``candidates``, ``dense_scores``, ``bm25_scores`` and optionally
        ``focused_scores`` (similarity with ``focus_query(...)``; omit or None when that text is empty) and
        ``path_entities``.  Sets ``final_score`` (:1629) and returns every query's candidates sorted by it."""
        nq = len(batch)
        lens = [len(b["candidates"]) for b in batch]
        nc = max(lens, default=0)
        if nc == 0:
            return [[] for _ in batch]
        dense = _padded([b["dense_scores"] for b in batch], lens, nc)
        bm25 = _padded([b["bm25_scores"] for b in batch], lens, nc)
        focused = _padded([[] if b.get("focused_scores") is None else b["focused_scores"] for b in batch], lens, nc)
        feat = np.zeros((5, nq, nc), np.float64)                     # cov, cons, prf similarity, focused weight, hop factor
        feat[4] = 1.0
        is_prf = np.zeros((nq, nc), np.uint8)
        group = np.zeros((nq, nc), np.int32)
        any_dup = False
        for r, b in enumerate(batch):
            wanted = set(e.lower() for e in (b.get("path_entities") or []))
            seen: Dict[Any, int] = {}
            for i, cand in enumerate(b["candidates"]):
                feat[0, r, i], feat[1, r, i] = self._cov_cons(cand, wanted)
                if cand.get("retrieval_method") == "prf_bridge":     # :1597-1601
                    is_prf[r, i] = 1
                    feat[2, r, i] = cand.get("similarity_score", 0.0)
                hop = int(cand.get("hop_no", 1))                     # :1621-1627
                feat[3, r, i] = self.focused_weight_hop2 if hop == 2 else max(0.0, self.focused_weight_hop2 - 0.10 * (hop - 2))
                if hop > 1 and 0.0 < self.hop_decay < 1.0:
                    feat[4, r, i] = self.hop_decay ** (hop - 1)
                g = seen.setdefault(cand.get("note_id"), i)          # candidates sharing a note id share one RRF entry
                group[r, i] = g
                any_dup |= g != i
        dev = self.device or torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):
            d_dense, d_bm25, d_foc = (torch.from_numpy(a).to(dev) for a in (dense, bm25, focused))
            d_feat, d_prf = torch.from_numpy(feat).to(dev), torch.from_numpy(is_prf).to(dev)
            d_group = torch.from_numpy(group).to(dev) if any_dup else None
            d_n = torch.tensor(lens, dtype=torch.int32, device=dev)
            out_s = torch.empty((nq, nc), dtype=torch.float64, device=dev)
            out_o = torch.empty((nq, nc), dtype=torch.int32, device=dev)
            nat.check(nat.lib.arb_rescore_khop(nq, nc, d_dense.data_ptr(), d_bm25.data_ptr(), d_foc.data_ptr(),
                                               d_feat[0].data_ptr(), d_feat[1].data_ptr(), d_feat[2].data_ptr(), d_prf.data_ptr(),
                                               d_group.data_ptr() if d_group is not None else None, d_feat[3].data_ptr(),
                                               d_feat[4].data_ptr(), d_n.data_ptr(), self.dense_weight, self.bm25_weight,
                                               self.cov_weight, self.cons_weight, self.rrf_lambda, self.RRF_K,
                                               out_s.data_ptr(), out_o.data_ptr(), _stream(dev)))
            score, order = out_s.cpu().numpy(), out_o.cpu().numpy()
        results = []
        for r, b in enumerate(batch):
            cands = b["candidates"]
            for i, cand in enumerate(cands):
                cand["final_score"] = float(score[r, i])
            results.append([cands[i] for i in order[r, :len(cands)].tolist()])
        return results

    def rerank_khop_candidates(self, candidates: List[Dict[str, Any]], dense_scores: Sequence[float],
                               bm25_scores: Sequence[float], focused_scores: Optional[Sequence[float]] = None,
                               path_entities: List[str] = None) -> List[Dict[str, Any]]:
        """Single-query form (:1549).  Unlike _enhanced_hybrid_search_v2 the reference does not catch here."""
        if not candidates:
            return []
        return self.rerank_batch([{"candidates": candidates, "dense_scores": dense_scores, "bm25_scores": bm25_scores,
                                   "focused_scores": focused_scores, "path_entities": path_entities}])[0]
