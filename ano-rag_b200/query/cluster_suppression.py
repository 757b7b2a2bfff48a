"""Near-duplicate cluster suppression of a re-ranked candidate list on the B200 -- drop-in for the reference's
``QueryProcessor._apply_cluster_suppression`` (query/query_processor.py:1637-1697).

The reference normalises every pair of candidate vectors and takes their dot product inside a Python double loop
(up to ~5 000 numpy calls for 100 candidates, per query).  Here the candidates' vectors of a whole batch of queries go
into one device index and ONE gather-scoring launch (``arb_index_score_rows``, csrc/dense.cu) returns every pairwise
cosine; the greedy walk over the resulting small boolean matrix, the two stable sorts and the config handling are the
reference's control flow and stay on the host.  Vectors come from ``_embed_vec`` or from the injected ``embed``
callable (the encoder is out of scope), exactly like :1661-1667.  Cosines within 1e-5 of the threshold may fall on
either side (fp32 rounding order); everywhere else the kept list is the reference's.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np

from .._log import logger
from .. import engine


def _parse(cs: Dict[str, Any]):
    enabled = cs.get("enabled", True)
    if isinstance(enabled, str):
        enabled = enabled.strip().lower() not in {"false", "0", "no"}                 # :1645-1648
    else:
        enabled = bool(enabled)
    try:
        thr = float(cs.get("cos_threshold", 0.90))
    except (TypeError, ValueError):
        thr = 0.90
    try:
        keep_m = int(cs.get("keep_per_cluster", 2))
    except (TypeError, ValueError):
        keep_m = 2
    return enabled, thr, keep_m


def pairwise_cosines(vec_lists: Sequence[np.ndarray]) -> List[np.ndarray]:
    """One [n_i, n_i] fp32 cosine matrix per list of candidate vectors; all lists share one index and one launch."""
    sizes = [len(v) for v in vec_lists]
    total = sum(sizes)
    if total == 0:
        return [np.zeros((0, 0), np.float32) for _ in vec_lists]
    dim = int(np.asarray(vec_lists[[i for i, n in enumerate(sizes) if n][0]][0]).shape[-1])
    rows = np.concatenate([np.asarray(v, dtype=np.float32).reshape(-1, dim) for v in vec_lists if len(v)])
    ix = engine.DenseIndex(dim, storage="fp32", normalize=True)                      # rows and queries divided by their norm
    try:
        ix.add(rows)
        width = max(sizes)
        cand = np.full((total, width), -1, dtype=np.int64)
        at = 0
        for n in sizes:
            cand[at:at + n, :n] = np.arange(at, at + n)
            at += n
        sims = ix.score_rows(rows, cand, clamp_nonneg=False)
    finally:
        ix.close()
    out, at = [], 0
    for n in sizes:
        out.append(sims[at:at + n, :n])
        at += n
    return out


def _suppress(candidates: List[Dict[str, Any]], cos: np.ndarray, thr: float, keep_m: int) -> List[Dict[str, Any]]:
    close = cos >= thr
    kept, used = [], np.zeros(len(candidates), dtype=bool)
    for i in range(len(candidates)):                                                  # :1672-1689
        if used[i]:
            continue
        members = [i] + [j for j in np.flatnonzero(close[i, i + 1:] & ~used[i + 1:]) + i + 1]
        members.sort(key=lambda k: candidates[k].get("final_score", 0.0), reverse=True)
        kept.extend(candidates[k] for k in members[:keep_m])
        used[members] = True
    kept.sort(key=lambda c: c.get("final_score", 0.0), reverse=True)                  # :1691
    return kept


def apply_cluster_suppression_batch(candidate_lists: Sequence[List[Dict[str, Any]]], cs_config: Optional[Dict[str, Any]] = None,
                                    embed: Optional[Callable[[str], Any]] = None) -> List[List[Dict[str, Any]]]:
    """The reference method for the candidate lists of several queries at once (one device launch for all cosines)."""
    enabled, thr, keep_m = _parse(cs_config or {})
    if not enabled:
        return [c for c in candidate_lists]
    vec_lists = []
    for cands in candidate_lists:
        vecs = []
        for c in cands:
            v = c.get("_embed_vec")
            if v is None:
                if embed is None:
                    raise ValueError("candidate without '_embed_vec' and no embed callable")
                v = embed(f"{c.get('title','')} {c.get('content','')}")                # :1664
                c["_embed_vec"] = v
            vecs.append(np.asarray(v, dtype=np.float32))
        vec_lists.append(vecs)
    cosines = pairwise_cosines(vec_lists)
    out = []
    for cands, cos in zip(candidate_lists, cosines):
        kept = _suppress(list(cands), cos, thr, keep_m) if cands else cands
        logger.info(f"Cluster suppression: {len(cands)} -> {len(kept)} candidates, threshold={thr}, keep_per_cluster={keep_m}")
        out.append(kept)
    return out


def apply_cluster_suppression(candidates: List[Dict[str, Any]], cs_config: Optional[Dict[str, Any]] = None,
                              embed: Optional[Callable[[str], Any]] = None) -> List[Dict[str, Any]]:
    """Signature-compatible single-query form (``candidates`` in, kept candidates out)."""
    if not candidates:
        return candidates
    return apply_cluster_suppression_batch([candidates], cs_config, embed)[0]
