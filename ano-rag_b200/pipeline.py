"""The hybrid candidate-scoring step as one call: dense top-k + BM25 top-k (+ the
cross-GPU exchange when the corpus is row-sharded) + fusion.

This is the batched form of what the reference executes per query in
QueryProcessor (vector_retriever.search -> bm25_scores -> fusion,
query/query_processor.py:1964,3706-3758) expressed over the same three components
(VectorIndex / bm25_search / HybridSearcher).

Sharding (new design): rank g holds rows [base_g, base_g + n_g) of
the embedding matrix and the postings of exactly those docs, with corpus-GLOBAL
BM25 statistics.  Per batch each rank computes its local top-k with global ids,
one all-to-all per list moves, to every rank, the candidates of ITS slice of the
queries (nq / world each) from all ranks over NVLink (NCCL), and the rank merges and fuses
that slice, so neither the merge work nor the exchanged bytes are replicated.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from . import engine
from .retrieval.hybrid_search import HybridSearcher


class PeerLists:
    """Home of one rank's local top-k lists in NVLink peer-mapped symmetric memory (torch symmetric memory is
    the allocator / rendezvous; the data path is the merge kernel's own loads from the peers' buffers).

    Two buffer sets alternate between batches: a rank starts writing set A for batch i+2 only after passing
    the barrier of batch i+1, which every rank reaches only after its merge of batch i (the last reader of
    set A) was enqueued before it on the same stream -- so one barrier per batch is enough."""

    _ALIGN = 256

    def __init__(self, group, nq: int, k_dense: int, k_bm25: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm
        self.nq, self.k_dense, self.k_bm25 = int(nq), int(k_dense), int(k_bm25)
        sizes = [nq * k_dense * 4, nq * k_dense * 8, nq * k_bm25 * 8, nq * k_bm25 * 8]
        self.offs, at = [], 0
        for b in sizes:
            self.offs.append(at)
            at += (b + self._ALIGN - 1) // self._ALIGN * self._ALIGN
        self.set_bytes = at
        group = group if group is not None else torch.distributed.group.WORLD
        with torch.cuda.device(device):
            self.buf = symm.empty(2 * self.set_bytes, dtype=torch.uint8, device=device)
            self.hdl = symm.rendezvous(self.buf, group)
        self.peer_ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.turn = 0

    def _view(self, which: int, dtype, k: int, nq: int) -> torch.Tensor:
        lo = self.turn * self.set_bytes + self.offs[which]
        return self.buf[lo: lo + nq * k * dtype.itemsize].view(dtype).view(nq, k)

    def dense_out(self, nq: int):
        return self._view(0, torch.float32, self.k_dense, nq), self._view(1, torch.int64, self.k_dense, nq)

    def bm25_out(self, nq: int):
        return self._view(2, torch.float64, self.k_bm25, nq), self._view(3, torch.int64, self.k_bm25, nq)

    def ptrs(self, which: int):
        return [p + self.turn * self.set_bytes + self.offs[which] for p in self.peer_ptrs]

    def barrier(self):
        self.hdl.barrier(channel=0)

    def flip(self):
        self.turn ^= 1


class HybridPipeline:
    def __init__(self, dense: engine.DenseIndex, bm25: Optional[engine.BM25Index], searcher: HybridSearcher,
                 k_dense: int = 100, k_bm25: int = 100, pool: Optional[int] = None, id_base: int = 0,
                 query_planes: Optional[int] = None, group=None, device: Optional[torch.device] = None):
        self.dense, self.bm25, self.searcher = dense, bm25, searcher
        self.k_dense, self.k_bm25 = int(k_dense), int(k_bm25)
        self.pool = int(pool if pool is not None else searcher.candidate_pool)
        self.id_base = int(id_base)
        self.query_planes = query_planes
        self.group = group
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.world = 1
        self.rank = 0
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(group)
            self.rank = torch.distributed.get_rank(group)
        # cross-GPU exchange: "peer" = merge kernel loads from peer-mapped symmetric memory (default, <= 16 ranks),
        # "nccl" = all-to-all of query slices followed by the merge
        self.exchange = os.environ.get("ARB_EXCHANGE", "peer" if self.world <= 16 else "nccl")
        self._peer: Optional[PeerLists] = None
        self._copy_stream = None
        self._bm25_stream = None
        self._bm25_out = None
        self._fused_ev = None
        self.overlap_max_queries = int(os.environ.get("ARB_OVERLAP_MAX_QUERIES", str(1 << 30)))
        self._host_out, self._host_turn = {}, {}

    # ------------------------------------------------------------------ device-resident inputs
    def _exchange(self, t: torch.Tensor, pad) -> torch.Tensor:
        """[nq, k] local lists -> [world, per, k]: from every rank, its lists for THIS rank's query slice.
        An all-to-all moves only the slice each rank merges (1/world of what an all-gather would move)."""
        nq, k = t.shape
        per = (nq + self.world - 1) // self.world
        send = t
        if per * self.world != nq:
            send = torch.full((per * self.world, k), pad, dtype=t.dtype, device=t.device)
            send[:nq] = t
        recv = torch.empty((self.world, per, k), dtype=t.dtype, device=t.device)
        torch.distributed.all_to_all_single(recv, send.contiguous(), group=self.group)
        return recv

    def _dense(self, q_emb: torch.Tensor, out, emb_ready):
        """Dense local top-k into ``out`` (or fresh tensors).  ``emb_ready`` = [(lo, hi, event), ...]: rows [lo, hi) of
        q_emb are valid once ``event`` (None = already) has happened -- lets the scan of the first rows of a batch
        start while the rest is still on its way from the host (search_host)."""
        nq = q_emb.shape[0]
        if out is None:
            out = (torch.empty((nq, self.k_dense), dtype=torch.float32, device=q_emb.device),
                   torch.empty((nq, self.k_dense), dtype=torch.int64, device=q_emb.device))
        for lo, hi, ev in (emb_ready or [(0, nq, None)]):
            if ev is not None:
                torch.cuda.current_stream(q_emb.device).wait_event(ev)
            if hi > lo:
                self.dense.search(q_emb[lo:hi], self.k_dense, query_planes=self.query_planes, id_base=self.id_base,
                                  out=(out[0][lo:hi], out[1][lo:hi]))
        return out

    def search_device(self, q_emb: torch.Tensor, q_terms: Optional[torch.Tensor], emb_ready=None, terms_ready=None):
        """q_emb [nq, d] CUDA (fp32 or bf16), q_terms [nq, max_qt] CUDA int32 (or None: dense only).
        -> (ids [nq,pool] int64, final [nq,pool] fp64, src [nq,pool,4] fp64, count [nq] int32) on the device.
        With world > 1 every rank returns the rows of its own query slice
        [rank*ceil(nq/world), ...) -- see ``query_slice``."""
        nq = q_emb.shape[0]
        with_bm25 = self.bm25 is not None and q_terms is not None
        if self.world > 1 and self.exchange == "peer":
            peer = self._peer_lists(nq)
            lo, hi = self.query_slice(nq)
            if with_bm25 and nq <= self.overlap_max_queries:      # the two sources on two streams (see below)
                main, side = torch.cuda.current_stream(q_emb.device), self._side()
                if terms_ready is not None:
                    side.wait_event(terms_ready)
                else:
                    side.wait_stream(main)
                with torch.cuda.stream(side):
                    self.bm25.search(q_terms, self.k_bm25, id_base=self.id_base, out=peer.bm25_out(nq))
                self._dense(q_emb, peer.dense_out(nq), emb_ready)
                main.wait_stream(side)
            else:
                self._dense(q_emb, peer.dense_out(nq), emb_ready)
                if with_bm25:
                    self.bm25.search(q_terms, self.k_bm25, id_base=self.id_base, out=peer.bm25_out(nq))
            peer.barrier()                       # every rank's lists are written and visible
            d_sc, d_id = engine.merge_topk_peers(peer.ptrs(0), peer.ptrs(1), torch.float32, lo, hi - lo,
                                                 self.k_dense, self.device)
            b_sc = b_id = None
            if with_bm25:
                b_sc, b_id = engine.merge_topk_peers(peer.ptrs(2), peer.ptrs(3), torch.float64, lo, hi - lo,
                                                     self.k_bm25, self.device)
            peer.flip()
            return self.searcher.fuse_tensors(dense=(d_id, d_sc), bm25=(b_id, b_sc) if with_bm25 else None,
                                              pool=self.pool)
        b_sc = b_id = None
        if with_bm25 and nq <= self.overlap_max_queries:
            # the two sources are independent until the fusion, so they run on two streams: small batches leave most of the
            # GPU idle in either stage (the dense scan streams the corpus, the BM25 stage waits on dependent loads) --
            # 0.70 -> 0.54 ms for one query, 1.71 -> 1.35 ms for 128; large batches gain the kernels' tails (~2 %)
            main, side = torch.cuda.current_stream(q_emb.device), self._side()
            if terms_ready is not None:
                side.wait_event(terms_ready)             # search_host: q_terms is on the device; the embeddings may still travel
                if self._fused_ev is not None:
                    side.wait_event(self._fused_ev)      # the previous batch's fusion has read the BM25 buffers
            else:
                side.wait_stream(main)
            # the BM25 lists live in buffers the pipeline keeps (written on the side stream, read by the fusion on the main
            # one): fresh tensors would have to be handed across streams with record_stream, which makes the caching
            # allocator hold their blocks back and allocate new ones -- a device-wide stall in the middle of a batch.
            # One set is enough: the next batch's BM25 starts only after the side stream has waited for this batch's fusion
            key = (nq, self.k_bm25, q_emb.device)
            if self._bm25_out is None or self._bm25_out[0] != key:
                self._bm25_out = (key, torch.empty((nq, self.k_bm25), dtype=torch.float64, device=q_emb.device),
                                  torch.empty((nq, self.k_bm25), dtype=torch.int64, device=q_emb.device))
                side.wait_stream(main)
            with torch.cuda.stream(side):
                b_sc, b_id = self.bm25.search(q_terms, self.k_bm25, id_base=self.id_base, out=self._bm25_out[1:])
            d_sc, d_id = self._dense(q_emb, None, emb_ready)
            main.wait_stream(side)
        else:
            d_sc, d_id = self._dense(q_emb, None, emb_ready)
            if with_bm25:
                b_sc, b_id = self.bm25.search(q_terms, self.k_bm25, id_base=self.id_base)
        if self.world > 1:
            lo, hi = self.query_slice(q_emb.shape[0])
            d_sc, d_id = engine.merge_topk(self._exchange(d_sc, float("-inf")), self._exchange(d_id, -1))
            d_sc, d_id = d_sc[: hi - lo], d_id[: hi - lo]
            if b_sc is not None:
                b_sc, b_id = engine.merge_topk(self._exchange(b_sc, 0.0), self._exchange(b_id, -1))
                b_sc, b_id = b_sc[: hi - lo], b_id[: hi - lo]
        res = self.searcher.fuse_tensors(dense=(d_id, d_sc), bm25=(b_id, b_sc) if b_sc is not None else None,
                                         pool=self.pool)
        if self._bm25_out is not None:
            if self._fused_ev is None:
                self._fused_ev = torch.cuda.Event()
            self._fused_ev.record(torch.cuda.current_stream(q_emb.device))
        return res

    def _side(self):
        if self._bm25_stream is None:
            self._bm25_stream = torch.cuda.Stream(device=self.device)
        return self._bm25_stream

    def _peer_lists(self, nq: int) -> PeerLists:
        """(Re)allocate the symmetric buffers; collective, and every rank sees the same nq."""
        if self._peer is None or self._peer.nq < nq:
            torch.cuda.synchronize(self.device)
            self._peer = PeerLists(self.group, nq, self.k_dense, self.k_bm25, self.device)
        return self._peer

    def query_slice(self, nq: int):
        if self.world == 1:
            return 0, nq
        per = (nq + self.world - 1) // self.world
        return min(nq, self.rank * per), min(nq, (self.rank + 1) * per)

    # ------------------------------------------------------------------ host buffers (the end-to-end call)
    def search_host(self, q_emb_host, q_terms_host, return_src: bool = False):
        """Host inputs -> host numpy outputs (ids, final, count[, src]); the H2D copy of the batch and the D2H copy of
        the fused result are part of the call.  With pinned inputs the second half of the embeddings travels on a
        copy stream while the dense scan of the first half runs; the results land in pinned buffers (one
        synchronisation).  The returned arrays are views of two alternating buffer sets: copy them if they must
        outlive the next call but one."""
        dev = self.device
        h_emb = torch.as_tensor(q_emb_host)
        nq = h_emb.shape[0]
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            q_emb = torch.empty(h_emb.shape, dtype=h_emb.dtype, device=dev)
            # the query terms (a few hundred KB) go first: the BM25 stream starts on them while the embeddings still travel
            q_terms = None
            terms_ready = None
            if q_terms_host is not None:
                q_terms = torch.as_tensor(q_terms_host).to(dev, non_blocking=True)
                terms_ready = torch.cuda.Event()
                terms_ready.record(main)
            half = ((nq // 2 + 127) // 128) * 128
            emb_ready = None
            if h_emb.is_pinned() and nq >= 1024 and 0 < half < nq:
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(device=dev)
                side = self._copy_stream
                side.wait_stream(main)                      # q_emb's memory may still be in use by earlier work
                q_emb[:half].copy_(h_emb[:half], non_blocking=True)
                ev = torch.cuda.Event()
                with torch.cuda.stream(side):
                    q_emb[half:].copy_(h_emb[half:], non_blocking=True)
                    ev.record(side)
                emb_ready = [(0, half, None), (half, nq, ev)]
            else:
                q_emb.copy_(h_emb, non_blocking=True)
            res = self.search_device(q_emb, q_terms, emb_ready=emb_ready, terms_ready=terms_ready)
            ids, fin, src, cnt = res
            outs = [ids, fin, cnt] + ([src] if return_src else [])
            key = tuple((tuple(t.shape), t.dtype) for t in outs)
            sets = self._host_out.get(key)
            if sets is None:
                sets = self._host_out[key] = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in outs] for _ in range(2)]
                self._host_turn[key] = 0
            turn = self._host_turn[key]
            self._host_turn[key] = turn ^ 1
            for buf, t in zip(sets[turn], outs):
                buf.copy_(t, non_blocking=True)
            main.synchronize()
            return tuple(buf.numpy() for buf in sets[turn])
