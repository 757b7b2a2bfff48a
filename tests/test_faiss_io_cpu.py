"""faiss on-disk index format (vector_store/faiss_io.py): byte layout against files assembled by hand from the
published serialiser layout, round trips, and the error cases.  faiss itself is absent (parity unpinned)."""
import struct

import numpy as np
import pytest

from ano_rag_b200.vector_store import faiss_io as fio


def _flat_bytes(rows, fourcc=b"IxFI", metric=0):
    n, d = rows.shape
    return (fourcc + struct.pack("<i", d) + struct.pack("<q", n) + struct.pack("<qq", 1 << 20, 1 << 20) + b"\x01"
            + struct.pack("<i", metric) + struct.pack("<Q", n * d) + rows.astype("<f4").tobytes())


def test_flat_layout_and_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    rows = rng.standard_normal((37, 12)).astype(np.float32)
    p = tmp_path / "a.faiss"
    fio.write_flat(str(p), rows)
    assert p.read_bytes() == _flat_bytes(rows)                 # 4 + 4 + 8 + 16 + 1 + 4 = 37-byte header, count, payload
    assert len(p.read_bytes()) == 37 + 8 + rows.nbytes
    for mmap in (True, False):
        got = fio.read_index(str(p), mmap=mmap)
        assert (got.kind, got.d, got.ntotal, got.metric, got.is_trained) == ("Flat", 12, 37, fio.METRIC_INNER_PRODUCT, True)
        assert np.array_equal(np.asarray(got.rows), rows)
    p.write_bytes(_flat_bytes(rows, b"IxF2", 1))
    assert fio.read_index(str(p)).metric == fio.METRIC_L2
    empty = np.zeros((0, 5), np.float32)
    fio.write_flat(str(p), empty)
    got = fio.read_index(str(p))
    assert got.ntotal == 0 and got.rows.shape == (0, 5)


@pytest.mark.parametrize("nlist,used", [(8, 8), (16, 3)])         # "full" and "sprs" size tables
def test_ivf_flat_layout_and_round_trip(tmp_path, nlist, used):
    rng = np.random.default_rng(nlist)
    n, d = 50, 6
    rows = rng.standard_normal((n, d)).astype(np.float32)
    cent = rng.standard_normal((nlist, d)).astype(np.float32)
    assign = rng.integers(0, used, n).astype(np.int32)
    p = tmp_path / "ivf.faiss"
    fio.write_ivf_flat(str(p), rows, cent, assign, nprobe=4)
    raw = p.read_bytes()
    # hand-assembled expectation
    want = b"IwFl" + struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, 0) + struct.pack("<QQ", nlist, 4)
    want += _flat_bytes(cent) + b"\x00" + struct.pack("<Q", 0) + b"ilar" + struct.pack("<QQ", nlist, 4 * d)
    sizes = np.bincount(assign, minlength=nlist)
    if (sizes > 0).sum() > nlist // 2:
        want += b"full" + struct.pack("<Q", nlist) + sizes.astype("<u8").tobytes()
    else:
        nz = np.nonzero(sizes)[0]
        want += b"sprs" + struct.pack("<Q", 2 * len(nz)) + b"".join(struct.pack("<QQ", int(i), int(sizes[i])) for i in nz)
    for li in range(nlist):
        ids = np.nonzero(assign == li)[0]
        if len(ids):
            want += rows[ids].astype("<f4").tobytes() + ids.astype("<i8").tobytes()
    assert raw == want
    got = fio.read_index(str(p))
    assert (got.kind, got.d, got.ntotal, got.nlist, got.nprobe) == ("IVFFlat", d, n, nlist, 4)
    assert np.array_equal(got.rows, rows) and np.array_equal(got.centroids, cent) and np.array_equal(got.assign, assign)


def test_reads_direct_map_variants_and_rejects_what_it_cannot_serve(tmp_path):
    rng = np.random.default_rng(5)
    n, d, nlist = 9, 4, 2
    rows = rng.standard_normal((n, d)).astype(np.float32)
    cent = rng.standard_normal((nlist, d)).astype(np.float32)
    assign = (np.arange(n) % 2).astype(np.int32)
    p = tmp_path / "x.faiss"
    fio.write_ivf_flat(str(p), rows, cent, assign, nprobe=1)
    raw = p.read_bytes()
    at = raw.index(b"ilar") - 9                                   # the direct map: type byte + empty vector
    assert raw[at: at + 9] == b"\x00" + struct.pack("<Q", 0)
    arr = struct.pack("<B", 1) + struct.pack("<Q", n) + np.arange(n, dtype="<i8").tobytes()         # DirectMap::Array
    p.write_bytes(raw[:at] + arr + raw[at + 9:])
    assert np.array_equal(fio.read_index(str(p)).rows, rows)
    ht = struct.pack("<B", 2) + struct.pack("<Q", 0) + struct.pack("<Q", 1) + struct.pack("<qq", 3, 7)   # Hashtable
    p.write_bytes(raw[:at] + ht + raw[at + 9:])
    assert np.array_equal(fio.read_index(str(p)).assign, assign)
    # ids that are not the running row numbers
    bad = bytearray(raw)
    bad[-8:] = struct.pack("<q", 100)
    p.write_bytes(bytes(bad))
    with pytest.raises(fio.FaissFormatError, match="permutation"):
        fio.read_index(str(p))
    p.write_bytes(raw[:-5])
    with pytest.raises(fio.FaissFormatError, match="truncated"):
        fio.read_index(str(p))
    p.write_bytes(b"IHNf" + raw[4:])
    with pytest.raises(fio.FaissFormatError, match="unsupported faiss index type"):
        fio.read_index(str(p))
    with pytest.raises(ValueError):
        fio.write_ivf_flat(str(p), rows, cent, assign + 5, nprobe=1)
