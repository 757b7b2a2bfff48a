"""Dense flat-IP top-k on the B200 against the oracle (through the C ABI).
Tolerance: ids identical except where |delta score| <= 1e-5 (north-star rule)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _mods():
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200 import engine
    from oracle import dense as od
    return engine, od


def make(n, d, nq, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = (x[rng.integers(0, n, nq)] + 0.5 * rng.standard_normal((nq, d))).astype(np.float32)
    return x, q


@pytest.mark.parametrize("n,d,nq,k", [
    (10_000, 384, 200, 20),      # BASELINE cfg1 shape
    (1, 64, 1, 1), (255, 64, 3, 10), (256, 64, 3, 10), (257, 128, 129, 10),
    (3000, 100, 17, 100),        # dim not a multiple of 64
    (5000, 768, 130, 256),       # max k, two query tiles
    (40, 32, 5, 100),            # k > N  -> -1 padding
    (70_001, 96, 300, 100),      # ragged last corpus tile, many work items
])
def test_fp32_exact_mode_matches_oracle(n, d, nq, k):
    engine, od = _mods()
    x, q = make(n, d, nq, seed=n + d)
    ix = engine.DenseIndex(d, storage="fp32", normalize=True)
    ix.add(x)
    assert ix.ntotal == n
    s, i = ix.search(q, k)
    ws, wi = od.flat_ip_topk(od.preprocess_vectors(q), od.preprocess_vectors(x), k)
    ok, n_match, total, worst = od.compare_topk(i, s, wi, ws, tol=TOL)
    assert ok, f"worst |ds|={worst}, id match {n_match}/{total}"
    assert n_match >= 0.95 * total
    assert np.all(np.diff(s, axis=1) <= 0)                      # descending
    if k > n:
        assert np.all(i[:, n:] == -1) and np.all(i[:, :n] >= 0)
    ix.close()


@pytest.mark.parametrize("n,d,nq,k", [(20_000, 768, 256, 100), (1000, 64, 5, 10), (33_333, 1024, 64, 100)])
def test_bf16_mode_exact_on_bf16_inputs(n, d, nq, k):
    """cfg2 recipe: rows normalised then rounded to bf16; products of bf16 values are exact in fp32,
    so only the summation order differs from the oracle."""
    import torch
    engine, od = _mods()
    x, q = make(n, d, nq, seed=7 * n)
    xb = od.bf16_round(od.preprocess_vectors(x))
    qb = od.bf16_round(od.preprocess_vectors(q))
    ix = engine.DenseIndex(d, storage="bf16", normalize=False)
    ix.add(torch.from_numpy(xb).to(torch.bfloat16).cuda())      # bf16 device tensor input
    s, i = ix.search(torch.from_numpy(qb).cuda(), k, query_planes=1)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    ws, wi = od.flat_ip_topk(qb, xb, k)
    ok, n_match, total, worst = od.compare_topk(i, s, wi, ws, tol=TOL)
    assert ok, f"worst |ds|={worst}, id match {n_match}/{total}"
    assert np.array_equal(ix.reconstruct(0, 50), xb[:50])       # storage holds exactly the bf16 values
    ix.close()


def test_bf16_storage_recall_vs_fp32_flat():
    engine, od = _mods()
    x, q = make(30_000, 384, 128, seed=5)
    ix = engine.DenseIndex(384, storage="bf16", normalize=True)
    ix.add(x)
    _, i = ix.search(q, 20, query_planes=1)
    _, wi = od.flat_ip_topk(od.preprocess_vectors(q), od.preprocess_vectors(x), 20)
    recall = np.mean([len(set(a) & set(b)) / 20 for a, b in zip(i.tolist(), wi.tolist())])
    assert recall >= 0.95, recall


def test_mixed_plane_modes_and_incremental_add():
    engine, od = _mods()
    x, q = make(9000, 200, 40, seed=3)
    xs, qs = od.preprocess_vectors(x), od.preprocess_vectors(q)
    ix = engine.DenseIndex(200, storage="fp32")
    for part in np.array_split(x, 7):                           # capacity growth path
        ix.add(part)
    s3, i3 = ix.search(q, 10, query_planes=3)
    s1, i1 = ix.search(q, 10, query_planes=1)                   # bf16 queries vs exact rows
    ws, wi = od.flat_ip_topk(qs, xs, 10)
    assert od.compare_topk(i3, s3, wi, ws, tol=TOL)[0]
    w1s, w1i = od.flat_ip_topk(od.bf16_round(qs), xs, 10)
    assert od.compare_topk(i1, s1, w1i, w1s, tol=TOL)[0]
    assert np.allclose(ix.reconstruct(100, 5), xs[100:105], atol=1e-6)
    ix.reset()
    assert ix.ntotal == 0
    ix.add(x[:10])
    assert ix.search(q[:2], 3)[1].max() < 10


def test_zero_rows_duplicates_and_ties():
    engine, od = _mods()
    rng = np.random.default_rng(0)
    base = rng.standard_normal((50, 64)).astype(np.float32)
    x = np.concatenate([base, base, np.zeros((6, 64), np.float32), base[:10]])      # exact duplicates + zero rows
    ix = engine.DenseIndex(64, storage="fp32")
    ix.add(x)
    q = np.concatenate([base[:8], np.zeros((1, 64), np.float32)])
    s, i = ix.search(q, 12)
    ws, wi = od.flat_ip_topk(od.preprocess_vectors(q), od.preprocess_vectors(x), 12)
    assert od.compare_topk(i, s, wi, ws, tol=TOL)[0]
    for r in range(8):                                          # the three copies of a row lead, ascending id
        assert set(i[r, :3]) == {r, r + 50, r + 106} and abs(s[r, 0] - 1.0) < 1e-5
        assert list(i[r, :3]) == sorted(i[r, :3]) or abs(s[r, 0] - s[r, 2]) <= TOL
    assert np.all(s[8] == 0.0)                                  # zero query (norm 0 -> 1)


def test_vector_index_dropin_matches_oracle_shaping(tmp_path):
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.config import config
    from ano_rag_b200.vector_store import VectorIndex
    from oracle import dense as od
    config.set("storage.vector_index_path", str(tmp_path))
    x, q = make(2000, 96, 6, seed=9)
    vi = VectorIndex(96)
    assert vi.index_type == "IVFFlat" and vi.nlist == 100 and vi.nprobe == 10
    assert vi.search(q) == []                                   # no index yet (:210-212)
    assert vi.create_index() and vi.add_vectors(x, np.arange(2000, dtype=np.int64))
    assert vi.total_vectors == 2000 and vi.is_trained
    got = vi.search(q, top_k=7)
    want = od.vector_index_search(q, x, top_k=7)
    assert len(got) == 6
    for g, w in zip(got, want):
        assert [r["rank"] for r in g] == [r["rank"] for r in w]
        assert [r["index"] for r in g] == [r["index"] for r in w]
        assert all(abs(a["score"] - b["score"]) <= TOL and a["similarity"] == a["score"] for a, b in zip(g, w))
    one = vi.search(q[:1], top_k=3)
    assert isinstance(one[0], dict) and len(one) == 3           # flat list for a single query (:256-257)
    # persistence round trip keeps results
    path = vi.save_index()
    assert path.endswith(".faiss")
    vi2 = VectorIndex(96)
    assert vi2.load_index(path.split("/")[-1]) and vi2.total_vectors == 2000
    again = vi2.search(q, top_k=7)
    assert [[r["index"] for r in g] for g in again] == [[r["index"] for r in g] for g in got]
    assert [[r["score"] for r in g] for g in again] == [[r["score"] for r in g] for g in got]   # rows restored bit for bit
    # the file is faiss's own IndexFlatIP format (what the reference's faiss.write_index leaves, :300)
    from ano_rag_b200.vector_store import faiss_io
    assert open(path, "rb").read(4) == b"IxFI"
    disk = faiss_io.read_index(path)
    assert disk.ntotal == 2000 and np.allclose(np.linalg.norm(np.asarray(disk.rows), axis=1), 1.0, atol=1e-6)
    assert np.allclose(np.asarray(disk.rows), od.preprocess_vectors(x), rtol=0, atol=3e-7)   # the reference's normalised rows
    # small training set shrinks nlist exactly like the reference (:140-155)
    vi3 = VectorIndex(96)
    vi3.create_index("IVFFlat")
    vi3.add_vectors(x[:50])
    assert vi3.nlist == 25
    stats = vi.get_index_stats()
    assert stats["ntotal"] == 2000 and stats["index_type"] == "IVFFlat" and stats["nprobe"] == 10
    before = vi.index.reconstruct(0, 2000)
    assert vi.remove_vectors(np.array([0, 1, 777])) and vi.index.ntotal == 1997
    keep = np.setdiff1d(np.arange(2000), [0, 1, 777])
    assert np.array_equal(vi.index.reconstruct(0, 1997), before[keep])       # survivors stay bit for bit (no second normalisation)
    res = vi.search(x[[5, 778]], top_k=1)
    assert res[0][0]["index"] == 3 and res[1][0]["index"] == 775              # faiss remove_ids semantics: later rows shift down
    vi.reset_index()
    assert vi.total_vectors == 0 and vi.search(q) == []
    config.reset()


def test_concurrent_searches_share_one_index():
    engine, od = _mods()
    x, q = make(20_000, 128, 64, seed=21)
    ix = engine.DenseIndex(128, storage="fp32")
    ix.add(x)
    ws, wi = ix.search(q, 10)
    out = [None] * 8

    def work(t):
        import torch
        with torch.cuda.stream(torch.cuda.Stream()):
            for _ in range(5):
                out[t] = ix.search(q, 10)

    th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    for s, i in out:
        assert np.array_equal(i, wi) and np.array_equal(s, ws)


def test_errors_are_reported_not_swallowed_at_the_abi():
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200 import engine, _native as nat
    ix = engine.DenseIndex(32)
    with pytest.raises(nat.NativeError):
        ix.search(np.zeros((1, 32), np.float32), 5)             # empty index
    ix.add(np.ones((4, 32), np.float32))
    with pytest.raises(nat.NativeError):
        ix.search(np.zeros((1, 32), np.float32), 1000)          # k beyond ARB_MAX_TOPK
    with pytest.raises(ValueError):
        ix.search(np.zeros((1, 31), np.float32), 5)


def test_prepass_and_cta_pair_variants_are_bit_identical(monkeypatch):
    """The two-pass search (pre-pass over a tenth of the rows seeds the thresholds) and the cta_group::2 kernel
    must return exactly what the plain single-CTA scan returns; the plain scan is pinned to the oracle above."""
    import torch
    engine, od = _mods()
    rng = np.random.default_rng(77)
    n, d, nq, k = 300_000, 64, 384, 50
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[1000:1040] = x[10:50]                                     # exact duplicates: ties across the pre-pass boundary
    q = (x[rng.integers(0, n, nq)] + 0.4 * rng.standard_normal((nq, d))).astype(np.float32)
    ix = engine.DenseIndex(d, storage="fp32", normalize=True)
    ix.add(x)
    qd = torch.from_numpy(q).cuda()
    out = {}
    for pair in ("0", "1"):
        for pre in ("0", "1"):
            monkeypatch.setenv("ARB_DENSE_PAIR", pair)
            monkeypatch.setenv("ARB_DENSE_PREPASS", pre)
            s, i = ix.search(qd, k)
            out[pair, pre] = (s.cpu().numpy(), i.cpu().numpy())
    base = out["0", "0"]
    for key, (s, i) in out.items():
        assert np.array_equal(i, base[1]) and np.array_equal(s, base[0]), key
    ws, wi = od.flat_ip_topk(od.preprocess_vectors(q[:16]), od.preprocess_vectors(x), k)
    ok, _, _, worst = od.compare_topk(base[1][:16], base[0][:16], wi, ws, tol=TOL)
    assert ok, worst
    ix.close()


def test_top_k_beyond_the_kernel_limit_returns_the_best_256():
    """faiss accepts any k; the select kernel keeps 256 candidates.  Asking for more returns the best 256 instead of an
    empty result (VectorRetriever.retrieve asks for 3 x top_k, vector_store/retriever.py:381)."""
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.config import config
    from ano_rag_b200.vector_store import VectorIndex
    from oracle import dense as od
    config.reset()
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2000, 48)).astype(np.float32)
    q = rng.standard_normal((3, 48)).astype(np.float32)
    vi = VectorIndex(embedding_dim=48)
    assert vi.create_index("Flat") and vi.add_vectors(x)
    res = vi.search(q, top_k=300)
    assert len(res) == 3 and all(len(r) == 256 for r in res)
    ws, wi = od.flat_ip_topk(od.preprocess_vectors(q), od.preprocess_vectors(x), 256)
    ok, *_ = od.compare_topk(np.array([[e["index"] for e in r] for r in res]), np.array([[e["score"] for e in r] for r in res]), wi, ws, tol=1e-5)
    assert ok
    vi.cleanup()


def test_add_vectors_keeps_custom_ids_like_add_with_ids():
    """vector_store/vector_index.py:187-193: non-Flat index types receive the caller's ids (add_with_ids); results carry
    them, remove_vectors takes them."""
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.config import config
    from ano_rag_b200.vector_store import VectorIndex
    config.reset()
    rng = np.random.default_rng(8)
    x = rng.standard_normal((500, 32)).astype(np.float32)
    vi = VectorIndex(embedding_dim=32)
    vi.index_type = "IVFFlat"
    assert vi.create_index("IVFFlat")
    ids = np.arange(500, dtype=np.int64) * 7 + 1000
    assert vi.add_vectors(x[:300], ids[:300]) and vi.add_vectors(x[300:], ids[300:])
    res = vi.search(x[[5, 410]], top_k=3)
    assert res[0][0]["index"] == 1035 and res[1][0]["index"] == 1000 + 7 * 410
    assert vi.remove_vectors(np.array([1035])) and vi.index.ntotal == 499
    res = vi.search(x[[5]], top_k=3)
    assert all(e["index"] != 1035 for e in res) and all((e["index"] - 1000) % 7 == 0 for e in res)
    assert abs(res[0]["score"] - float(np.sort((x / np.linalg.norm(x, axis=1, keepdims=True)) @ (x[5] / np.linalg.norm(x[5])))[-2])) < 1e-5
    vi.cleanup()
