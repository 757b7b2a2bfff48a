"""Parity at the BASELINE.json scales (cfg2 / cfg3: 1 M notes x 768-d bf16, 4096 queries, top-100; cfg5 shape at 1 M rows).

The oracle cannot score 4096 queries against 1 M notes in seconds, so a seeded sample of queries is checked in full
(ordered ids, bit-exact BM25 and fused scores, dense under the north star's 1e-5 tie rule) and the whole batch through
size-independent properties: descending order, positive BM25 scores only, the batch result equal to the same queries run
in small batches (another kernel path), fused lists equal to the fusion of the per-source lists.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, D, NQ, K, V = 1_000_000, 768, 4096, 100, 50_000


@pytest.fixture(scope="module")
def cfg3():
    import torch
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200 import bm25_build as bb, engine, synth
    from ano_rag_b200.pipeline import HybridPipeline
    from ano_rag_b200.retrieval.hybrid_search import HybridSearcher
    dev = torch.device("cuda:0")
    dense = engine.DenseIndex(D, storage="bf16", normalize=False)
    for _, rows in synth.dense_rows(0, N, D, dev, host_rng=True):
        dense.add(rows)
    q = synth.dense_queries(NQ, D, dev, n_rows=N, host_rng=True)
    ptr, tok = synth.bm25_corpus(N, V)
    qt = synth.bm25_queries(NQ, ptr, tok, V)
    arr = bb.build_arrays(ptr, tok, V, variant="okapi")
    bm = engine.BM25Index(N, arr.term_ptr, arr.post_doc, arr.post_gcode, arr.stats.idf, arr.gtab)
    cfg = {"retrieval": {"candidate_pool": K, "hybrid": {"fusion_method": "rrf", "rrf_k": 60, "weights": {"dense": 1.0, "bm25": 1.0}}}}
    pipe = HybridPipeline(dense, bm, HybridSearcher(cfg), k_dense=K, k_bm25=K, pool=K, query_planes=1, device=dev)
    yield {"dense": dense, "bm": bm, "pipe": pipe, "q": q, "qt": qt, "ptr": ptr, "tok": tok, "cfg": cfg, "dev": dev}
    dense.close(); bm.close()


def test_cfg2_dense_top100_of_one_million(cfg3):
    """BASELINE configs[1]: exact IP top-100, bf16 rows, against the fp32 flat-IP oracle on the same rounded values."""
    import torch
    from ano_rag_b200 import synth
    from oracle import dense as od
    s, i = cfg3["dense"].search(cfg3["q"], K, query_planes=1)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    assert np.all(np.diff(s, axis=1) <= 0) and i.min() >= 0 and i.max() < N
    assert all(len(set(r)) == K for r in i[::97])
    rng = np.random.default_rng(11)
    sample = np.sort(rng.choice(NQ, 48, replace=False))
    x = np.concatenate([r.to(torch.float32).numpy() for _, r in synth.dense_rows(0, N, D, "cpu", host_rng=True)])
    ws, wi = od.flat_ip_topk(cfg3["q"][sample].to(torch.float32).cpu().numpy(), x, K)
    ok, n_match, total, worst = od.compare_topk(i[sample], s[sample], wi, ws, tol=1e-5)
    assert ok, f"dense top-100 differs from the oracle beyond 1e-5 ties (worst gap {worst})"
    assert n_match / total >= 0.95
    # the same queries in batches of 128 (one query tile, another split plan) give the same ids
    s2, i2 = cfg3["dense"].search(cfg3["q"][:128].contiguous(), K, query_planes=1)
    assert np.array_equal(i2.cpu().numpy(), i[:128]) and np.array_equal(s2.cpu().numpy(), s[:128])


def test_cfg3_bm25_and_fused_ranking_of_one_million(cfg3):
    """BASELINE configs[2]: BM25 top-100 bit-exact and the fused ORDERED (id, score) pairs exact on sampled queries;
    the whole batch equals the per-query kernel on a larger sample."""
    import torch
    from oracle import bm25 as obm, fusion as ofu
    qt_d = torch.from_numpy(cfg3["qt"]).to(cfg3["dev"])
    bs, bi = cfg3["bm"].search(qt_d, K)
    bs, bi = bs.cpu().numpy(), bi.cpu().numpy()
    valid = bi >= 0
    assert np.all(bs[valid] > 0) and np.all(bs[~valid] == 0) and np.all(np.diff(bs, axis=1) <= 0)
    ids, fin, src, cnt = (t.cpu().numpy() for t in cfg3["pipe"].search_device(cfg3["q"], qt_d))
    assert np.all(np.diff(fin, axis=1)[np.arange(K - 1)[None, :] < (cnt[:, None] - 1)] <= 0)
    ds, di = cfg3["dense"].search(cfg3["q"], K, query_planes=1)
    ds, di = ds.cpu().numpy(), di.cpu().numpy()
    ora = obm.ArrayBM25Oracle(cfg3["ptr"], cfg3["tok"], variant="okapi")
    fuser = ofu.HybridSearcherOracle(cfg3["cfg"])
    rng = np.random.default_rng(12)
    for r in np.sort(rng.choice(NQ, 64, replace=False)):
        ws, wi = obm.topk_positive(ora.get_scores([t for t in cfg3["qt"][r] if t >= 0]), K)
        assert np.array_equal(bi[r], wi) and np.array_equal(bs[r], ws), f"BM25 top-100 of query {r}"
        want = fuser.fuse([(int(a), float(b)) for a, b in zip(di[r], ds[r]) if a >= 0], [(int(a), float(b)) for a, b in zip(wi, ws) if a >= 0])
        got = [(int(ids[r, j]), float(fin[r, j])) for j in range(int(cnt[r]))]
        assert got == [(e["note_id"], e["final_similarity"]) for e in want], f"fused ranking of query {r}"
    # tile-major batched pipeline == one CTA per query, on 200 queries
    os.environ["ARB_BM25_MODE"] = "query"
    try:
        s0, i0 = cfg3["bm"].search(qt_d[:200].contiguous(), K)
    finally:
        del os.environ["ARB_BM25_MODE"]
    assert np.array_equal(i0.cpu().numpy(), bi[:200]) and np.array_equal(s0.cpu().numpy(), bs[:200])


def test_cfg5_shape_ivf_recall_at_one_million_rows():
    """BASELINE configs[4] in shape (topic-mixture rows, centroids = per-topic means, nprobe 32) at 1 M x 128-d, 1024 lists:
    recall@100 of the IVF search against the exact flat search of the same index, and nprobe = nlist equal to flat."""
    import torch
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200 import engine
    dev = torch.device("cuda:0")
    n, d, nlist, nq, nprobe = 1_000_000, 128, 1024, 2048, 32
    g = torch.Generator(device=dev); g.manual_seed(5)
    centres = torch.nn.functional.normalize(torch.randn((nlist, d), generator=g, device=dev), dim=1)
    topic = torch.randint(0, nlist, (n,), generator=g, device=dev)
    x = torch.nn.functional.normalize(centres[topic] + 0.3 * torch.randn((n, d), generator=g, device=dev) / d ** 0.5, dim=1)
    sums = torch.zeros((nlist, d), device=dev).index_add_(0, topic, x)
    cent = sums / torch.bincount(topic, minlength=nlist).clamp(min=1)[:, None]           # doc/clustering.py:194: the cluster mean
    q = torch.nn.functional.normalize(x[torch.randint(0, n, (nq,), generator=g, device=dev)] + 0.2 * torch.randn((nq, d), generator=g, device=dev) / d ** 0.5, dim=1)
    ix = engine.DenseIndex(d, storage="bf16", normalize=False)
    ix.set_centroids(cent.contiguous())
    ix.add(x.to(torch.bfloat16))
    fs, fi = ix.search(q.to(torch.bfloat16), 100, query_planes=1)
    vs, vi = ix.ivf_search(q.to(torch.bfloat16), 100, nprobe, query_planes=1)
    fi, vi = fi.cpu().numpy(), vi.cpu().numpy()
    recall = np.mean([len(set(a) & set(b)) / 100.0 for a, b in zip(fi, vi)])
    print(f"cfg5-shape recall@100 (nprobe {nprobe} of {nlist}) = {recall:.4f}")
    assert recall >= 0.95
    assert np.all(np.diff(vs.cpu().numpy(), axis=1) <= 0)
    ix.close()
