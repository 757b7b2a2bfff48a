"""Reader / writer for the two faiss on-disk index formats the reference produces on this path.

The reference persists its index with ``faiss.write_index`` (``vector_store/vector_index.py:300``) and loads it
with ``faiss.read_index`` (``:335``); ``VectorRetriever`` looks for ``*.faiss`` next to ``atomic_notes.json``
(``vector_store/retriever.py:689-706``).  To be a drop-in on an existing work directory the B200 index reads
and writes the same files for the index types that are served natively:

* ``IndexFlatIP`` / ``IndexFlatL2``  (fourcc ``IxFI`` / ``IxF2``)  -- created at ``vector_index.py:76-80``
* ``IndexIVFFlat``                   (fourcc ``IwFl``)            -- created at ``vector_index.py:84-86``, with
  ``ArrayInvertedLists`` (``ilar``, list sizes stored ``full`` or ``sprs``) and a flat coarse quantiser

faiss itself (``faiss-gpu-cu12 >= 1.7.4``, requirements.txt:8-10) is a third-party dependency that is absent from
the reference tree and from this image, so the layout below is a restatement of its published serialiser
(``faiss/impl/index_write.cpp`` / ``index_read.cpp``: ``write_index_header``, ``write_ivf_header``,
``write_direct_map``, ``write_InvertedLists``) -- PARITY UNPINNED: no faiss-written fixture could be generated
here; the tests pin the byte layout against a hand-assembled file and round trips.

Layout (little endian; ``idx_t`` = int64, ``size_t`` = uint64; a "vector" is a uint64 element count followed
by the raw elements):

    header   := d:int32  ntotal:int64  dummy:int64 (1<<20)  dummy:int64 (1<<20)  is_trained:uint8
                metric_type:int32 (0 = inner product, 1 = L2)  [metric_arg:float32 if metric_type > 1]
    IxFI/IxF2 := fourcc  header  vector<float32>[ntotal * d]
    IwFl     := fourcc  header  nlist:uint64  nprobe:uint64  <quantiser index>  direct_map  invlists
    direct_map := type:uint8 (0 = none, 1 = array, 2 = hashtable)  vector<int64>  [hashtable: vector<(int64,int64)>]
    invlists := "ilar"  nlist:uint64  code_size:uint64  ("full" vector<uint64>[nlist] | "sprs" vector<uint64>[2 * n_non_empty])
                then for every non-empty list: codes (n * code_size bytes)  ids (n * int64)
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import BinaryIO, Optional

import numpy as np

METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1
_DUMMY = 1 << 20


class FaissFormatError(ValueError):
    pass


@dataclass
class FaissIndexData:
    kind: str                       # 'Flat' | 'IVFFlat'
    d: int
    ntotal: int
    metric: int                     # METRIC_INNER_PRODUCT | METRIC_L2
    is_trained: bool
    rows: np.ndarray                # fp32 [ntotal, d] in id order (may be a read-only memmap for Flat)
    nlist: int = 0
    nprobe: int = 1
    centroids: Optional[np.ndarray] = None     # fp32 [nlist, d]
    assign: Optional[np.ndarray] = None        # int32 [ntotal] list of every row (IVFFlat)


# ----------------------------------------------------------------------------- reading
def _read(f: BinaryIO, fmt: str):
    size = struct.calcsize(fmt)
    raw = f.read(size)
    if len(raw) != size:
        raise FaissFormatError("truncated faiss index file")
    return struct.unpack(fmt, raw)


def _read_header(f: BinaryIO):
    d, ntotal, _, _, trained, metric = _read(f, "<iqqqBi")
    if metric > 1:
        _read(f, "<f")
    if d <= 0 or ntotal < 0:
        raise FaissFormatError(f"implausible faiss header: d={d} ntotal={ntotal}")
    return d, ntotal, bool(trained), metric


def _read_vector(f: BinaryIO, dtype, path: Optional[str] = None, mmap: bool = False) -> np.ndarray:
    (n,) = _read(f, "<Q")
    dt = np.dtype(dtype)
    if mmap and path is not None and n:
        off = f.tell()
        f.seek(n * dt.itemsize, 1)
        return np.memmap(path, dtype=dt, mode="r", offset=off, shape=(n,))
    arr = np.fromfile(f, dtype=dt, count=n)
    if arr.size != n:
        raise FaissFormatError("truncated faiss index file")
    return arr


def _read_flat(f: BinaryIO, fourcc: bytes, path: Optional[str], mmap: bool) -> FaissIndexData:
    d, ntotal, trained, metric = _read_header(f)
    xb = _read_vector(f, np.float32, path, mmap)
    if xb.size != ntotal * d:
        raise FaissFormatError(f"IndexFlat payload has {xb.size} floats, header says {ntotal} x {d}")
    return FaissIndexData("Flat", d, ntotal, metric, trained, xb.reshape(ntotal, d))


def read_index(path: str, mmap: bool = True) -> FaissIndexData:
    """Parse ``path``; Flat payloads are memory-mapped (they are copied to the device in chunks anyway)."""
    with open(path, "rb") as f:
        fourcc = f.read(4)
        if fourcc in (b"IxFI", b"IxF2"):
            return _read_flat(f, fourcc, path, mmap)
        if fourcc != b"IwFl":
            raise FaissFormatError(f"unsupported faiss index type {fourcc!r} (served natively: IxFI, IxF2, IwFl)")
        d, ntotal, trained, metric = _read_header(f)
        nlist, nprobe = _read(f, "<QQ")
        qcc = f.read(4)
        if qcc not in (b"IxFI", b"IxF2"):
            raise FaissFormatError(f"unsupported coarse quantiser {qcc!r}")
        quant = _read_flat(f, qcc, None, False)
        if quant.d != d or (trained and quant.ntotal != nlist):
            raise FaissFormatError("coarse quantiser does not match the IVF header")
        (dm_type,) = _read(f, "<B")
        _read_vector(f, np.int64)                                   # direct map array: derived data, rebuilt on demand
        if dm_type == 2:
            _read_vector(f, np.dtype([("k", "<i8"), ("v", "<i8")]))
        il = f.read(4)
        rows = np.zeros((ntotal, d), np.float32)
        assign = np.full(ntotal, -1, np.int32)
        if il == b"il00":
            if ntotal:
                raise FaissFormatError("IVF index without inverted lists but ntotal > 0")
        elif il == b"ilar":
            il_nlist, code_size = _read(f, "<QQ")
            if il_nlist != nlist or code_size != 4 * d:
                raise FaissFormatError(f"inverted lists: nlist={il_nlist} code_size={code_size}, expected {nlist} / {4 * d}")
            kind = f.read(4)
            sizes = np.zeros(nlist, np.int64)
            raw = _read_vector(f, np.uint64)
            if kind == b"full":
                if raw.size != nlist:
                    raise FaissFormatError("inverted lists: size table length != nlist")
                sizes[:] = raw
            elif kind == b"sprs":
                pairs = raw.reshape(-1, 2).astype(np.int64)
                sizes[pairs[:, 0]] = pairs[:, 1]
            else:
                raise FaissFormatError(f"inverted lists: unknown size table {kind!r}")
            if int(sizes.sum()) != ntotal:
                raise FaissFormatError("inverted lists hold a different number of vectors than ntotal")
            seen = np.zeros(ntotal, bool)
            for li in np.nonzero(sizes)[0]:
                n = int(sizes[li])
                codes = np.fromfile(f, dtype=np.float32, count=n * d)
                ids = np.fromfile(f, dtype=np.int64, count=n)
                if codes.size != n * d or ids.size != n:
                    raise FaissFormatError("truncated faiss index file")
                if ids.min() < 0 or ids.max() >= ntotal or seen[ids].any():
                    raise FaissFormatError("IVF ids are not a permutation of 0..ntotal-1 (the reference adds arange ids, "
                                           "vector_store/retriever.py:156); sparse ids are not supported")
                seen[ids] = True
                rows[ids] = codes.reshape(n, d)
                assign[ids] = li
        else:
            raise FaissFormatError(f"unsupported inverted list container {il!r}")
        return FaissIndexData("IVFFlat", d, ntotal, metric, trained, rows, int(nlist), int(nprobe),
                              np.ascontiguousarray(quant.rows) if quant.ntotal else None, assign)


# ----------------------------------------------------------------------------- writing
def _write_header(f: BinaryIO, d: int, ntotal: int, trained: bool, metric: int) -> None:
    f.write(struct.pack("<iqqqBi", d, ntotal, _DUMMY, _DUMMY, 1 if trained else 0, metric))


def _write_flat(f: BinaryIO, rows: np.ndarray, d: int, metric: int) -> None:
    f.write(b"IxFI" if metric == METRIC_INNER_PRODUCT else b"IxF2")
    _write_header(f, d, len(rows), True, metric)
    f.write(struct.pack("<Q", rows.size))
    np.ascontiguousarray(rows, dtype="<f4").tofile(f)


def write_flat(path: str, rows: np.ndarray, metric: int = METRIC_INNER_PRODUCT) -> None:
    rows = np.asarray(rows, dtype=np.float32)
    with open(path, "wb") as f:
        _write_flat(f, rows, rows.shape[1], metric)


def write_ivf_flat(path: str, rows: np.ndarray, centroids: np.ndarray, assign: np.ndarray, nprobe: int,
                   metric: int = METRIC_INNER_PRODUCT) -> None:
    """rows fp32 [n, d] in id order, centroids fp32 [nlist, d], assign[i] = list of row i."""
    rows = np.asarray(rows, dtype=np.float32)
    centroids = np.asarray(centroids, dtype=np.float32)
    assign = np.asarray(assign, dtype=np.int64)
    n, d = rows.shape
    nlist = len(centroids)
    if assign.shape != (n,) or (n and (assign.min() < 0 or assign.max() >= nlist)):
        raise ValueError("assign must hold one list number in [0, nlist) per row")
    order = np.argsort(assign, kind="stable")                       # ids ascending inside each list, like sequential adds
    sizes = np.bincount(assign, minlength=nlist).astype(np.uint64)
    with open(path, "wb") as f:
        f.write(b"IwFl")
        _write_header(f, d, n, True, metric)
        f.write(struct.pack("<QQ", nlist, nprobe))
        _write_flat(f, centroids, d, metric)
        f.write(struct.pack("<BQ", 0, 0))                           # DirectMap::NoMap, empty array
        f.write(b"ilar")
        f.write(struct.pack("<QQ", nlist, 4 * d))
        non_empty = np.nonzero(sizes)[0]
        if len(non_empty) > nlist // 2:
            f.write(b"full")
            f.write(struct.pack("<Q", nlist))
            sizes.astype("<u8").tofile(f)
        else:
            f.write(b"sprs")
            f.write(struct.pack("<Q", 2 * len(non_empty)))
            np.stack([non_empty.astype(np.uint64), sizes[non_empty]], axis=1).astype("<u8").tofile(f)
        at = 0
        for li in non_empty:
            ids = order[at: at + int(sizes[li])]
            at += int(sizes[li])
            np.ascontiguousarray(rows[ids], dtype="<f4").tofile(f)
            ids.astype("<i8").tofile(f)
