"""Row-sharding helpers (new design; the reference has no distributed path).

Rank g of G holds the contiguous rows [lo_g, hi_g) of the embedding matrix and the postings of
exactly those docs.  BM25 statistics must be corpus-GLOBAL or shard scores differ from the
single-index scores, so df, the token count and the vocabulary's first-appearance order are
combined with three small all-reduces (any backend: NCCL tensors on the GPU box, gloo on CPU).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import bm25_build as bb


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of ``rank``: [lo, hi)."""
    per = (n + world - 1) // world
    return min(n, rank * per), min(n, (rank + 1) * per)


def grid_layout(world: int, rank: int, n_shards: int, n_rows: int, n_queries: int):
    """``world`` GPUs as ``n_shards`` row shards x ``world / n_shards`` query groups (``bench.py --row-shards``).

    Rank r holds row shard ``r % n_shards`` and serves the query slice of group ``r // n_shards``; the ranks
    ``[g * n_shards, (g + 1) * n_shards)`` of one group exchange and merge their candidates, each finishing 1 / n_shards of
    the group's queries.  -> dict(shard, group, rows=(lo, hi), queries=(lo, hi), peers=[ranks of the group],
    out_queries=(lo, hi): the GLOBAL query range whose fused results this rank returns)."""
    if n_shards < 1 or world % n_shards:
        raise ValueError(f"{n_shards} row shards do not divide {world} GPUs")
    n_groups = world // n_shards
    shard, grp = rank % n_shards, rank // n_shards
    r_lo, r_hi = shard_range(n_rows, n_shards, shard)
    q_lo, q_hi = shard_range(n_queries, n_groups, grp)
    o_lo, o_hi = shard_range(q_hi - q_lo, n_shards, shard) if n_shards > 1 else (0, q_hi - q_lo)
    return {"shard": shard, "group": grp, "n_groups": n_groups, "rows": (r_lo, r_hi), "queries": (q_lo, q_hi),
            "peers": list(range(grp * n_shards, (grp + 1) * n_shards)), "out_queries": (q_lo + o_lo, q_lo + o_hi)}


def global_bm25_stats(doc_ptr: np.ndarray, doc_tok: np.ndarray, n_terms: int, token_base: int,
                      variant: str = "okapi", k1: float = bb.K1_DEFAULT, b: float = bb.B_DEFAULT,
                      epsilon: float = bb.EPS_DEFAULT, group=None, device: Optional[torch.device] = None) -> bb.BM25Stats:
    """Corpus-global BM25 statistics from this rank's docs.

    ``token_base`` is the global position of this shard's first token (sum of the token counts of
    the lower ranks), needed only to reproduce rank_bm25's vocabulary order for ``average_idf``.
    """
    doc_ptr = np.asarray(doc_ptr, dtype=np.int64)
    doc_tok = np.asarray(doc_tok, dtype=np.int64)
    term_ptr, _, _ = bb.term_doc_counts(doc_ptr, doc_tok, n_terms)
    df = np.diff(term_ptr)
    first = np.full(n_terms, np.iinfo(np.int64).max, dtype=np.int64)
    if len(doc_tok):
        np.minimum.at(first, doc_tok, token_base + np.arange(len(doc_tok), dtype=np.int64))
    counts = np.array([len(doc_ptr) - 1, len(doc_tok)], dtype=np.int64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dev = device or torch.device("cpu")
        t_df = torch.from_numpy(df.copy()).to(dev)
        t_first = torch.from_numpy(first).to(dev)
        t_counts = torch.from_numpy(counts).to(dev)
        dist.all_reduce(t_df, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(t_first, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(t_counts, op=dist.ReduceOp.SUM, group=group)
        df, first, counts = t_df.cpu().numpy(), t_first.cpu().numpy(), t_counts.cpu().numpy()
    present = np.flatnonzero(df > 0)
    appearance = present[np.argsort(first[present], kind="stable")].tolist()
    return bb.compute_stats(variant, int(counts[0]), int(counts[1]), df, appearance, k1, b, epsilon)


def token_bases(n_tokens_local: int, group=None, device: Optional[torch.device] = None) -> int:
    """Exclusive prefix sum of the per-rank token counts (this rank's ``token_base``)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = device or torch.device("cpu")
    mine = torch.zeros(world, dtype=torch.int64, device=dev)
    mine[rank] = n_tokens_local
    dist.all_reduce(mine, op=dist.ReduceOp.SUM, group=group)
    return int(mine[:rank].sum().item())
