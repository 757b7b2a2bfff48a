"""IVF (IndexIVFFlat replacement) on the B200 against the restated faiss contract (oracle/dense.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def clustered(n, d, n_centres, nq, seed):
    """BASELINE cfg5 recipe in small: unit-norm topic centres + noise; centroids = per-cluster means."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((n_centres, d)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    topic = rng.integers(0, n_centres, n)
    x = centres[topic] + 0.3 * rng.standard_normal((n, d)).astype(np.float32) / np.sqrt(d)
    x = (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    cent = np.stack([x[topic == c].mean(0) if np.any(topic == c) else centres[c] for c in range(n_centres)]).astype(np.float32)
    q = x[rng.integers(0, n, nq)] + 0.2 * rng.standard_normal((nq, d)).astype(np.float32) / np.sqrt(d)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    return x, cent, q


def check_against_oracle(ix, x, q, k, nprobe):
    from oracle import dense as od
    cent, assign = ix.ivf_export()
    # the quantiser is the same exact inner-product search: assignments agree outside fp32 near-ties
    want_assign = od.ivf_assign(x, cent)
    margin_ok = assign == want_assign
    if not margin_ok.all():
        s = x[~margin_ok] @ cent.T
        top2 = np.sort(s, axis=1)[:, -2:]
        assert np.all(top2[:, 1] - top2[:, 0] <= TOL), "assignment differs outside a near-tie"
    gs, gi = ix.ivf_search(q, k, nprobe)
    ws, wi, probes = od.ivf_search(q, x, cent, assign, k, nprobe)
    ok, n_match, total, worst = od.compare_topk(gi, gs, wi, ws, tol=TOL)
    if not ok:
        # a probed-list choice may legitimately differ when two centroids tie within TOL for the nprobe-th place
        cs = np.sort(q @ cent.T, axis=1)[:, ::-1]
        near = np.abs(cs[:, nprobe - 1] - cs[:, min(nprobe, cent.shape[0] - 1)]) <= TOL
        bad = [r for r in range(len(q)) if not near[r] and not od.compare_topk(gi[r:r + 1], gs[r:r + 1], wi[r:r + 1], ws[r:r + 1], tol=TOL)[0]]
        assert not bad, f"queries {bad[:5]} differ from the oracle (worst |dscore| {worst})"
    return gs, gi


@pytest.mark.parametrize("storage,n,d,nlist,nq,k,nprobe", [("fp32", 20_000, 96, 64, 300, 10, 4), ("fp32", 5_000, 64, 7, 33, 100, 7),
                                                          ("bf16", 60_000, 128, 256, 1000, 50, 16), ("fp32", 3_000, 40, 300, 5, 256, 40)])
def test_ivf_given_centroids_matches_oracle(storage, n, d, nlist, nq, k, nprobe):
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200 import engine
    from oracle import dense as od
    x, cent, q = clustered(n, d, nlist, nq, seed=n + d)
    if storage == "bf16":
        x, cent, q = od.bf16_round(x), od.bf16_round(cent), od.bf16_round(q)
    ix = engine.DenseIndex(d, storage=storage, normalize=False)
    ix.add(x[: n // 2])                       # rows added before the centroids exist are quantised on install
    ix.set_centroids(cent)
    ix.add(x[n // 2:])
    assert ix.nlist == nlist and ix.ntotal == n
    gs, gi = check_against_oracle(ix, x, q, k, nprobe)
    # probing every list is the exact flat search
    fs, fi = ix.search(q, min(k, 100))
    as_, ai = ix.ivf_search(q, min(k, 100), nlist)
    ok, *_ = od.compare_topk(ai, as_, fi, fs, tol=TOL)
    assert ok
    # device tensors in -> device tensors out, same numbers
    import torch
    ts, ti = ix.ivf_search(torch.from_numpy(q).cuda(), k, nprobe)
    assert np.array_equal(ti.cpu().numpy(), gi) and np.array_equal(ts.cpu().numpy(), gs)
    ix.close()


def test_ivf_training_and_recall():
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200 import engine
    from oracle import dense as od
    x, _, q = clustered(40_000, 64, 50, 200, seed=5)
    ix = engine.DenseIndex(64, storage="fp32", normalize=True)
    ix.train(x[:20_000], nlist=64, niter=10)
    ix.add(x)
    cent, assign = ix.ivf_export()
    assert cent.shape == (64, 64) and np.allclose(np.linalg.norm(cent, axis=1), 1.0, atol=1e-4)   # spherical k-means
    assert np.bincount(assign, minlength=64).max() < 40_000 // 4                                     # no degenerate clustering
    check_against_oracle(ix, od.preprocess_vectors(x), od.preprocess_vectors(q), 10, 8)
    fs, fi = ix.search(q, 10)
    rec = []
    for nprobe in (1, 4, 16, 64):
        _, ii = ix.ivf_search(q, 10, nprobe)
        rec.append(np.mean([len(set(a) & set(b)) / 10 for a, b in zip(ii, fi)]))
    assert rec == sorted(rec) and rec[-1] == 1.0 and rec[1] > 0.8, rec
    ix.close()


def test_vector_index_ivf_lists_mode(tmp_path):
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.config import config
    from ano_rag_b200.vector_store import VectorIndex
    config.set("storage.vector_index_path", str(tmp_path))
    config.set("b200.ivf", "lists")
    config.set("vector_store.nlist", 32)
    config.set("vector_store.nprobe", 32)
    x, cent, q = clustered(8_000, 48, 32, 20, seed=3)
    vi = VectorIndex(48)
    assert vi.create_index() and not vi.is_trained
    assert vi.add_vectors(x) and vi.is_trained and vi.index.nlist == 32      # auto-train (:177-180)
    exact = vi.index.search(q, 5)[1]
    got = vi.search(q, top_k=5)
    assert [[r["index"] for r in g] for g in got] == exact.tolist()          # nprobe == nlist: exact
    truth = exact
    best = vi.optimize_search_params(q, truth, target_recall=0.9)            # sweep of :437-462
    assert best["nprobe"] in (1, 5, 10, 20) and best["recall"] >= 0.9 and vi.nprobe == best["nprobe"]
    path = vi.save_index()
    assert open(path, "rb").read(4) == b"IwFl"                               # faiss IndexIVFFlat container
    from ano_rag_b200.vector_store import faiss_io
    disk = faiss_io.read_index(path)
    assert disk.kind == "IVFFlat" and disk.nlist == 32 and np.array_equal(disk.assign, vi.index.ivf_export()[1])
    vi2 = VectorIndex(48)
    assert vi2.load_index(path.split("/")[-1]) and vi2.index.nlist == 32
    assert np.array_equal(vi2.index.ivf_export()[1], disk.assign)            # same lists after the device re-assignment
    vi2.nprobe = vi.nprobe
    assert [[r["index"] for r in g] for g in vi2.search(q, top_k=5)] == [[r["index"] for r in g] for g in vi.search(q, top_k=5)]
    vi3 = VectorIndex(48)
    vi3.create_index()
    assert vi3.set_centroids(cent) and vi3.add_vectors(x) and vi3.nlist == 32
    config.reset()
