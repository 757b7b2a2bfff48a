"""Tensor-level host API over the C ABI (include/anorag_b200.h).

These classes hold opaque library handles and pass raw pointers + sizes + the
current CUDA stream; PyTorch is used only to own device memory and streams.
numpy arrays are treated as HOST buffers (the library stages them and synchronises),
torch CUDA tensors as DEVICE buffers (asynchronous on the current stream).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native as nat

try:  # torch is plumbing: device memory + streams
    import torch
except Exception:  # pragma: no cover - torch is part of the image
    torch = None


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def _stream_ptr(device=None) -> int:
    if torch is None or not torch.cuda.is_available():
        return 0
    return int(torch.cuda.current_stream(device).cuda_stream)


class _Buf:
    """Pointer view of a numpy array (host) or torch tensor (host or device)."""

    def __init__(self, x, dtypes_np, dtypes_torch, name):
        self.keep = x
        if _is_torch(x):
            if x.dtype not in dtypes_torch:
                raise TypeError(f"{name}: unsupported torch dtype {x.dtype}")
            if not x.is_contiguous():
                x = x.contiguous()
                self.keep = x
            self.ptr = x.data_ptr()
            self.is_device = 1 if x.is_cuda else 0
            self.device = x.device if x.is_cuda else None
            self.code = {torch.float32: nat.ARB_F32, torch.bfloat16: nat.ARB_BF16, torch.float64: nat.ARB_F64}.get(x.dtype, -1)
            self.shape = tuple(x.shape)
        else:
            x = np.asarray(x)
            if x.dtype not in dtypes_np:
                raise TypeError(f"{name}: unsupported numpy dtype {x.dtype}")
            if not x.flags["C_CONTIGUOUS"]:
                x = np.ascontiguousarray(x)
            self.keep = x
            self.ptr = x.ctypes.data
            self.is_device = 0
            self.device = None
            self.code = {np.dtype(np.float32): nat.ARB_F32, np.dtype(np.float64): nat.ARB_F64}.get(x.dtype, -1)
            self.shape = tuple(x.shape)


class DenseIndex:
    """Exact inner-product index (the IndexFlatIP replacement).

    storage='fp32' keeps three bf16 planes per row (exact split of the fp32 value,
    searched with six tensor-core passes, |score error| ~1e-7 for unit rows);
    storage='bf16' keeps bf16(row) (one pass; the BASELINE cfg2/cfg4 layout).
    """

    def __init__(self, dim: int, storage: str = "fp32", normalize: bool = True):
        if storage not in ("fp32", "bf16"):
            raise ValueError("storage must be 'fp32' or 'bf16'")
        self.dim = int(dim)
        self.storage = storage
        self.normalize = bool(normalize)
        h = ctypes.c_void_p()
        nat.check(nat.lib.arb_index_create(ctypes.byref(h), self.dim,
                                           nat.STORE_FP32_SPLIT if storage == "fp32" else nat.STORE_BF16,
                                           1 if normalize else 0))
        self._h = h

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            nat.lib.arb_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def ntotal(self) -> int:
        return int(nat.lib.arb_index_ntotal(self._h))

    def reset(self):
        nat.check(nat.lib.arb_index_reset(self._h))

    def add_prepared(self, rows) -> None:
        """Append rows that are already pre-processed (a saved index holds them L2-normalised): stored as they are."""
        nat.check(nat.lib.arb_index_set_row_normalization(self._h, 0))
        try:
            self.add(rows)
        finally:
            nat.check(nat.lib.arb_index_set_row_normalization(self._h, 1 if self.normalize else 0))

    def add(self, rows) -> None:
        f32 = (np.dtype(np.float32),)
        b = _Buf(rows, f32, (torch.float32, torch.bfloat16) if torch is not None else (), "rows")
        if len(b.shape) != 2 or b.shape[1] != self.dim:
            raise ValueError(f"rows must be [n, {self.dim}], got {b.shape}")
        nat.check(nat.lib.arb_index_add(self._h, b.ptr, b.code, b.shape[0], b.is_device, _stream_ptr(b.device)))

    def reconstruct(self, row0: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = self.ntotal - row0 if n is None else n
        out = np.empty((n, self.dim), dtype=np.float32)
        nat.check(nat.lib.arb_index_reconstruct(self._h, row0, n, out.ctypes.data, 0, _stream_ptr()))
        return out

    def reconstruct_device(self, row0: int = 0, n: Optional[int] = None):
        """The stored rows as an fp32 CUDA tensor [n, dim] (no host round trip)."""
        n = self.ntotal - row0 if n is None else n
        dev = torch.device("cuda", torch.cuda.current_device())
        out = torch.empty((n, self.dim), dtype=torch.float32, device=dev)
        if n:
            nat.check(nat.lib.arb_index_reconstruct(self._h, row0, n, out.data_ptr(), 1, _stream_ptr(dev)))
        return out

    def search(self, queries, k: int, query_planes: Optional[int] = None, id_base: int = 0, out=None):
        """-> (scores fp32 [nq,k], ids int64 [nq,k]); numpy in -> numpy out, CUDA tensor in -> CUDA tensors out.
        out = (scores, ids): contiguous CUDA tensors to write into (device queries only)."""
        if query_planes is None:
            query_planes = 3 if self.storage == "fp32" else 1
        b = _Buf(queries, (np.dtype(np.float32),), (torch.float32, torch.bfloat16) if torch is not None else (), "queries")
        if len(b.shape) != 2 or b.shape[1] != self.dim:
            raise ValueError(f"queries must be [nq, {self.dim}], got {b.shape}")
        nq = b.shape[0]
        if b.is_device:
            with torch.cuda.device(b.device):
                scores, ids = _out_lists(out, nq, k, torch.float32, b.device)
                nat.check(nat.lib.arb_index_search(self._h, b.ptr, b.code, nq, 1, k, query_planes, id_base,
                                                   scores.data_ptr(), ids.data_ptr(), 1, _stream_ptr(b.device)))
            return scores, ids
        scores = np.empty((nq, k), dtype=np.float32)
        ids = np.empty((nq, k), dtype=np.int64)
        nat.check(nat.lib.arb_index_search(self._h, b.ptr, b.code, nq, 0, k, query_planes, id_base,
                                           scores.ctypes.data, ids.ctypes.data, 0, _stream_ptr()))
        return scores, ids


    def score_rows(self, queries, cand_rows, query_planes: Optional[int] = None, clamp_nonneg: bool = True) -> np.ndarray:
        """Similarity of query q with stored row cand_rows[q, c] -> fp32 [nq, nc] (host); rows < 0 score 0."""
        if query_planes is None:
            query_planes = 3 if self.storage == "fp32" else 1
        b = _Buf(queries, (np.dtype(np.float32),), (torch.float32, torch.bfloat16) if torch is not None else (), "queries")
        cand = np.ascontiguousarray(cand_rows, dtype=np.int64)
        if len(b.shape) != 2 or b.shape[1] != self.dim or cand.ndim != 2 or cand.shape[0] != b.shape[0]:
            raise ValueError("queries must be [nq, dim] and cand_rows [nq, nc]")
        out = np.empty(cand.shape, dtype=np.float32)
        nat.check(nat.lib.arb_index_score_rows(self._h, b.ptr, b.code, b.shape[0], b.is_device, cand.ctypes.data, cand.shape[1], 0,
                                               query_planes, 1 if clamp_nonneg else 0, out.ctypes.data, 0, _stream_ptr(b.device)))
        return out

    # -- IVF (IndexIVFFlat): the same handle, once it has centroids
    @property
    def nlist(self) -> int:
        return int(nat.lib.arb_index_ivf_nlist(self._h))

    def set_centroids(self, centroids) -> None:
        b = _Buf(centroids, (np.dtype(np.float32),), (torch.float32,) if torch is not None else (), "centroids")
        if len(b.shape) != 2 or b.shape[1] != self.dim:
            raise ValueError(f"centroids must be [nlist, {self.dim}], got {b.shape}")
        nat.check(nat.lib.arb_index_ivf_set_centroids(self._h, b.shape[0], b.ptr, b.is_device, _stream_ptr(b.device)))

    def train(self, rows, nlist: int, niter: int = 25, spherical: bool = True) -> None:
        b = _Buf(rows, (np.dtype(np.float32),), (torch.float32, torch.bfloat16) if torch is not None else (), "rows")
        if len(b.shape) != 2 or b.shape[1] != self.dim:
            raise ValueError(f"rows must be [n, {self.dim}], got {b.shape}")
        nat.check(nat.lib.arb_index_ivf_train(self._h, int(nlist), b.ptr, b.code, b.shape[0], b.is_device, int(niter),
                                              1 if spherical else 0, _stream_ptr(b.device)))

    def ivf_export(self):
        """-> (centroids fp32 [nlist, dim], list of every row int32 [ntotal])"""
        cent = np.empty((self.nlist, self.dim), dtype=np.float32)
        assign = np.empty(self.ntotal, dtype=np.int32)
        nat.check(nat.lib.arb_index_ivf_export(self._h, cent.ctypes.data, assign.ctypes.data, _stream_ptr()))
        return cent, assign

    def ivf_search(self, queries, k: int, nprobe: int, query_planes: Optional[int] = None, id_base: int = 0):
        """Same contract as search(), restricted to the nprobe lists nearest to each query."""
        if query_planes is None:
            query_planes = 3 if self.storage == "fp32" else 1
        b = _Buf(queries, (np.dtype(np.float32),), (torch.float32, torch.bfloat16) if torch is not None else (), "queries")
        if len(b.shape) != 2 or b.shape[1] != self.dim:
            raise ValueError(f"queries must be [nq, {self.dim}], got {b.shape}")
        nq = b.shape[0]
        if b.is_device:
            with torch.cuda.device(b.device):
                scores = torch.empty((nq, k), dtype=torch.float32, device=b.device)
                ids = torch.empty((nq, k), dtype=torch.int64, device=b.device)
                nat.check(nat.lib.arb_index_ivf_search(self._h, b.ptr, b.code, nq, 1, k, int(nprobe), query_planes, id_base,
                                                       scores.data_ptr(), ids.data_ptr(), 1, _stream_ptr(b.device)))
            return scores, ids
        scores = np.empty((nq, k), dtype=np.float32)
        ids = np.empty((nq, k), dtype=np.int64)
        nat.check(nat.lib.arb_index_ivf_search(self._h, b.ptr, b.code, nq, 0, k, int(nprobe), query_planes, id_base,
                                               scores.ctypes.data, ids.ctypes.data, 0, _stream_ptr()))
        return scores, ids


class BM25Index:
    """Device-resident block-partitioned inverted index (see csrc/bm25.cu)."""

    def __init__(self, n_docs: int, term_ptr: np.ndarray, post_doc: np.ndarray, post_gcode: np.ndarray,
                 idf: np.ndarray, gtab: np.ndarray):
        term_ptr = np.ascontiguousarray(term_ptr, dtype=np.int64)
        post_doc = np.ascontiguousarray(post_doc, dtype=np.int32)
        post_gcode = np.ascontiguousarray(post_gcode, dtype=np.int32)
        idf = np.ascontiguousarray(idf, dtype=np.float64)
        gtab = np.ascontiguousarray(gtab, dtype=np.float64)
        self.n_docs = int(n_docs)
        self.n_terms = len(term_ptr) - 1
        if len(idf) != self.n_terms or len(post_doc) != term_ptr[-1] or len(post_gcode) != len(post_doc):
            raise ValueError("inconsistent CSR arrays")
        h = ctypes.c_void_p()
        nat.check(nat.lib.arb_bm25_create(ctypes.byref(h), self.n_docs, self.n_terms, term_ptr.ctypes.data,
                                          post_doc.ctypes.data, post_gcode.ctypes.data, idf.ctypes.data,
                                          gtab.ctypes.data, len(gtab), _stream_ptr()))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            nat.lib.arb_bm25_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_postings(self) -> int:
        return int(nat.lib.arb_bm25_npostings(self._h))

    def query_postings(self, q_terms: np.ndarray) -> int:
        q = np.ascontiguousarray(q_terms, dtype=np.int32)
        return int(nat.lib.arb_bm25_query_postings(self._h, q.ctypes.data, q.shape[0], q.shape[1]))

    @staticmethod
    def _terms(q_terms):
        if _is_torch(q_terms):
            b = _Buf(q_terms, (), (torch.int32,), "q_terms")
        else:
            b = _Buf(np.ascontiguousarray(q_terms, dtype=np.int32), (np.dtype(np.int32),), (), "q_terms")
        if len(b.shape) != 2:
            raise ValueError("q_terms must be [nq, max_qt]")
        return b

    def search(self, q_terms, k: int, id_base: int = 0, out=None):
        """-> (scores fp64 [nq,k], ids int64 [nq,k]) descending, -1 padded.
        out = (scores, ids): contiguous CUDA tensors to write into (device queries only)."""
        b = self._terms(q_terms)
        nq, mq = b.shape
        if b.is_device:
            with torch.cuda.device(b.device):
                scores, ids = _out_lists(out, nq, k, torch.float64, b.device)
                nat.check(nat.lib.arb_bm25_search(self._h, b.ptr, nq, mq, 1, k, id_base, scores.data_ptr(),
                                                  ids.data_ptr(), 1, _stream_ptr(b.device)))
            return scores, ids
        scores = np.empty((nq, k), dtype=np.float64)
        ids = np.empty((nq, k), dtype=np.int64)
        nat.check(nat.lib.arb_bm25_search(self._h, b.ptr, nq, mq, 0, k, id_base, scores.ctypes.data,
                                          ids.ctypes.data, 0, _stream_ptr()))
        return scores, ids

    def scores_device(self, q_terms, n_out: Optional[int] = None, normalise: bool = False, continue_from=None):
        """Full fp64 score vectors [nq, n_out] left on the device (torch CUDA tensor).  ``continue_from``: the tensor an
        earlier call returned for the FIRST tokens of the same queries -- the sums continue in it, token by token, so a
        query longer than the kernel's per-call limit keeps the reference's summation order."""
        b = self._terms(q_terms)
        nq, mq = b.shape
        n_out = self.n_docs if n_out is None else int(n_out)
        device = b.device if b.is_device else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(device):
            out = continue_from if continue_from is not None else torch.empty((nq, n_out), dtype=torch.float64, device=device)
            if tuple(out.shape) != (nq, n_out) or out.dtype != torch.float64 or not out.is_contiguous():
                raise ValueError("continue_from must be the contiguous fp64 [nq, n_out] tensor of the earlier call")
            flags = (1 if normalise else 0) | (2 if continue_from is not None else 0)
            nat.check(nat.lib.arb_bm25_scores(self._h, b.ptr, nq, mq, b.is_device, flags, n_out,
                                              out.data_ptr(), 1, _stream_ptr(device)))
        return out

    def scores(self, q_terms, n_out: Optional[int] = None, normalise: bool = False) -> np.ndarray:
        """Full fp64 score vectors [nq, n_out] on the host (bm25_scores semantics)."""
        b = self._terms(q_terms)
        nq, mq = b.shape
        n_out = self.n_docs if n_out is None else int(n_out)
        out = np.empty((nq, n_out), dtype=np.float64)
        if b.is_device:
            with torch.cuda.device(b.device):
                nat.check(nat.lib.arb_bm25_scores(self._h, b.ptr, nq, mq, 1, 1 if normalise else 0, n_out,
                                                  out.ctypes.data, 0, _stream_ptr(b.device)))
        else:
            nat.check(nat.lib.arb_bm25_scores(self._h, b.ptr, nq, mq, 0, 1 if normalise else 0, n_out,
                                              out.ctypes.data, 0, _stream_ptr()))
        return out


def combine_fields(field_scores: Sequence, weights: Sequence[float], normalise_rows: bool):
    """FieldWeightedBM25 combination of per-field CUDA score tensors [nq, n] -> CUDA tensor [nq, n]."""
    first = field_scores[0]
    nq, n = first.shape
    out = torch.empty_like(first)
    ptrs = (ctypes.c_void_p * len(field_scores))(*[t.data_ptr() for t in field_scores])
    w = (ctypes.c_double * len(weights))(*[float(x) for x in weights])
    with torch.cuda.device(first.device):
        nat.check(nat.lib.arb_combine_fields(len(field_scores), ptrs, w, nq * n, nq if normalise_rows else 0,
                                             out.data_ptr(), _stream_ptr(first.device)))
    return out


_SRC = ("dense", "bm25", "graph", "path")


def fuse_batch(method: str, weights, rrf_k: float, pool: int, dense=None, bm25=None, graph=None, path=None):
    """Batched HybridSearcher.fuse.  Each source is ``(ids int64 [nq,k_s], scores [nq,k_s])`` or None;
    all numpy (host) or all torch CUDA tensors.  ``weights`` is a mapping or 4-sequence in
    dense/bm25/graph/path order.  -> (ids [nq,pool], final [nq,pool], src [nq,pool,4], count [nq])."""
    srcs = [dense, bm25, graph, path]
    if isinstance(weights, dict):
        w = [float(weights.get(k, 0.0)) for k in _SRC]
    else:
        w = [float(x) for x in weights]
    present = [s for s in srcs if s is not None]
    if not present:
        raise ValueError("fuse_batch needs at least one source")
    on_device = _is_torch(present[0][0]) and present[0][0].is_cuda
    nq = int(present[0][0].shape[0])
    ids_p = (ctypes.c_void_p * 4)()
    sc_p = (ctypes.c_void_p * 4)()
    dt = (ctypes.c_int * 4)()
    ks = (ctypes.c_int * 4)()
    keep = []
    device = None
    for i, s in enumerate(srcs):
        if s is None:
            ids_p[i] = None; sc_p[i] = None; dt[i] = nat.ARB_F64; ks[i] = 0
            continue
        ids, sc = s
        if on_device:
            bi = _Buf(ids, (), (torch.int64,), "ids")
            bs = _Buf(sc, (), (torch.float32, torch.float64), "scores")
            device = bi.device
        else:
            bi = _Buf(np.ascontiguousarray(ids, dtype=np.int64), (np.dtype(np.int64),), (), "ids")
            sc = np.asarray(sc)
            if sc.dtype != np.float32:
                sc = sc.astype(np.float64, copy=False)
            bs = _Buf(sc, (np.dtype(np.float32), np.dtype(np.float64)), (), "scores")
        if bi.shape != bs.shape or len(bi.shape) != 2 or bi.shape[0] != nq:
            raise ValueError("source ids/scores must both be [nq, k]")
        keep += [bi, bs]
        ids_p[i] = bi.ptr; sc_p[i] = bs.ptr; dt[i] = bs.code; ks[i] = bi.shape[1]
    wv = (ctypes.c_double * 4)(*w)
    code = nat.FUSE_RRF if method == "rrf" else nat.FUSE_LINEAR
    if on_device:
        with torch.cuda.device(device):
            o_ids = torch.empty((nq, pool), dtype=torch.int64, device=device)
            o_fin = torch.empty((nq, pool), dtype=torch.float64, device=device)
            o_src = torch.empty((nq, pool, 4), dtype=torch.float64, device=device)
            o_cnt = torch.empty((nq,), dtype=torch.int32, device=device)
            nat.check(nat.lib.arb_fuse(code, wv, float(rrf_k), pool, nq, ids_p, sc_p, dt, ks, o_ids.data_ptr(),
                                       o_fin.data_ptr(), o_src.data_ptr(), o_cnt.data_ptr(), 1, _stream_ptr(device)))
        return o_ids, o_fin, o_src, o_cnt
    o_ids = np.empty((nq, pool), dtype=np.int64)
    o_fin = np.empty((nq, pool), dtype=np.float64)
    o_src = np.empty((nq, pool, 4), dtype=np.float64)
    o_cnt = np.empty((nq,), dtype=np.int32)
    nat.check(nat.lib.arb_fuse(code, wv, float(rrf_k), pool, nq, ids_p, sc_p, dt, ks, o_ids.ctypes.data,
                               o_fin.ctypes.data, o_src.ctypes.data, o_cnt.ctypes.data, 0, _stream_ptr()))
    return o_ids, o_fin, o_src, o_cnt


def _out_lists(out, nq, k, score_dtype, device):
    if out is None:
        return (torch.empty((nq, k), dtype=score_dtype, device=device),
                torch.empty((nq, k), dtype=torch.int64, device=device))
    scores, ids = out
    for t, dt in ((scores, score_dtype), (ids, torch.int64)):
        if not (t.is_cuda and t.device == device and t.dtype == dt and tuple(t.shape) == (nq, k) and t.is_contiguous()):
            raise ValueError(f"out tensors must be contiguous CUDA [{nq}, {k}] {score_dtype} / int64 on {device}")
    return scores, ids


def merge_topk_peers(score_ptrs, id_ptrs, score_dtype, row0: int, nq: int, k: int, device):
    """Merge rows [row0, row0 + nq) of `parts` per-GPU [rows, k] lists given by raw (peer-mapped) device
    addresses -> ([nq,k], [nq,k]) on `device`; see arb_merge_topk_peers."""
    import ctypes
    parts = len(score_ptrs)
    code = nat.ARB_F32 if score_dtype == torch.float32 else nat.ARB_F64
    sp = (ctypes.c_void_p * parts)(*score_ptrs)
    ip = (ctypes.c_void_p * parts)(*id_ptrs)
    with torch.cuda.device(device):
        o_s = torch.empty((nq, k), dtype=score_dtype, device=device)
        o_i = torch.empty((nq, k), dtype=torch.int64, device=device)
        nat.check(nat.lib.arb_merge_topk_peers(code, sp, ip, parts, row0, nq, k, o_s.data_ptr(), o_i.data_ptr(),
                                               _stream_ptr(device)))
    return o_s, o_i


def merge_topk(scores, ids):
    """scores/ids CUDA tensors [parts, nq, k] (fp32 or fp64 / int64) -> merged ([nq,k], [nq,k])."""
    bs = _Buf(scores, (), (torch.float32, torch.float64), "scores")
    bi = _Buf(ids, (), (torch.int64,), "ids")
    if not bs.is_device or bs.shape != bi.shape or len(bs.shape) != 3:
        raise ValueError("merge_topk expects CUDA tensors [parts, nq, k]")
    parts, nq, k = bs.shape
    with torch.cuda.device(bs.device):
        o_s = torch.empty((nq, k), dtype=bs.keep.dtype, device=bs.device)
        o_i = torch.empty((nq, k), dtype=torch.int64, device=bs.device)
        nat.check(nat.lib.arb_merge_topk(bs.code, bs.ptr, bi.ptr, parts, nq, k, 0, o_s.data_ptr(), o_i.data_ptr(),
                                         _stream_ptr(bs.device)))
    return o_s, o_i


class TfidfIndex:
    """Device-resident L2-normalised TF-IDF document matrix, term-major (see csrc/tfidf.cu): the cosine top-k of
    the reference's "BM25 fallback" (vector_store/retriever.py:966-1002)."""

    def __init__(self, matrix):
        """``matrix``: scipy sparse [n_docs, n_terms] with rows already L2-normalised (sklearn ``normalize``)."""
        csc = matrix.tocsc()
        csc.sort_indices()
        self.n_docs, self.n_terms = int(csc.shape[0]), int(csc.shape[1])
        ptr = np.ascontiguousarray(csc.indptr, dtype=np.int64)
        doc = np.ascontiguousarray(csc.indices, dtype=np.int32)
        val = np.ascontiguousarray(csc.data, dtype=np.float64)
        h = ctypes.c_void_p()
        nat.check(nat.lib.arb_tfidf_create(ctypes.byref(h), self.n_docs, self.n_terms, ptr.ctypes.data, doc.ctypes.data,
                                           val.ctypes.data))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            nat.lib.arb_tfidf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def search(self, queries, k: int):
        """``queries``: scipy sparse [nq, n_terms], rows L2-normalised -> (scores fp64 [nq,k], ids int64 [nq,k]);
        descending similarity, larger doc id first among equal ones, only similarities > 0 (id -1 beyond)."""
        csr = queries.tocsr()
        csr.sort_indices()
        nq = int(csr.shape[0])
        ptr = np.ascontiguousarray(csr.indptr, dtype=np.int64)
        term = np.ascontiguousarray(csr.indices, dtype=np.int32)
        val = np.ascontiguousarray(csr.data, dtype=np.float64)
        scores = np.empty((nq, k), dtype=np.float64)
        ids = np.empty((nq, k), dtype=np.int64)
        nat.check(nat.lib.arb_tfidf_search(self._h, nq, ptr.ctypes.data, term.ctypes.data, val.ctypes.data, int(k),
                                           scores.ctypes.data, ids.ctypes.data, _stream_ptr()))
        return scores, ids
