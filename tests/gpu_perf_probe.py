"""Quick timing at BASELINE sizes (not the bench; a development probe)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ano_rag_b200
from ano_rag_b200 import engine, bm25_build as bb

what = sys.argv[1:] or ["dense", "bm25"]
dev = torch.device("cuda:0")

def timeit(fn, warm=2, it=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(it + 1)]
    ev[0].record()
    for i in range(it):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(it)]

if "dense" in what:
    N, D, Q, K = 1_000_000, 768, 4096, 100
    g = torch.Generator(device=dev); g.manual_seed(1234)
    ix = engine.DenseIndex(D, storage="bf16", normalize=False)
    for s in range(0, N, 250_000):
        x = torch.randn((250_000, D), generator=g, device=dev)
        x = torch.nn.functional.normalize(x, dim=1).to(torch.bfloat16)
        ix.add(x)
    q = torch.nn.functional.normalize(torch.randn((Q, D), generator=g, device=dev), dim=1).to(torch.bfloat16)
    ts = timeit(lambda: ix.search(q, K, query_planes=1))
    ms = min(ts)
    print(f"dense 1Mx768 bf16 Q=4096 k=100: {ts} ms -> {Q/ms*1e3:.0f} q/s, {2*Q*N*D/ms/1e9:.1f} TFLOP/s", flush=True)
    # spot check vs torch on a slice of queries
    s, i = ix.search(q[:64], K, query_planes=1)
    xr = torch.from_numpy(ix.reconstruct(0, 200_000)).to(dev)
    print("check subset vs torch skipped (full corpus needed)", flush=True)

if "bm25" in what:
    N, V, Q, K = 1_000_000, 50_000, 4096, 100
    rng = np.random.default_rng(1236)
    t0 = time.time()
    lens = np.maximum(1, rng.poisson(40, N))
    ptr = np.zeros(N + 1, np.int64); ptr[1:] = np.cumsum(lens)
    p = 1.0 / np.arange(1, V + 1); p /= p.sum()
    cdf = np.cumsum(p)
    tok = np.searchsorted(cdf, rng.random(int(ptr[-1]))).astype(np.int64).clip(0, V - 1)
    print(f"gen {time.time()-t0:.1f}s tokens={len(tok)}", flush=True)
    t0 = time.time()
    arr = bb.build_arrays(ptr, tok, V, variant="okapi")
    print(f"build arrays {time.time()-t0:.1f}s postings={len(arr.post_doc)} n_g={len(arr.gtab)}", flush=True)
    t0 = time.time()
    ix = engine.BM25Index(N, arr.term_ptr, arr.post_doc, arr.post_gcode, arr.stats.idf, arr.gtab)
    print(f"device index {time.time()-t0:.1f}s", flush=True)
    qt = np.empty((Q, 8), np.int32)
    qt[:, :4] = np.searchsorted(cdf, rng.random((Q, 4))).clip(0, V - 1)
    docs = rng.integers(0, N, Q)
    for r in range(Q):
        d = tok[ptr[docs[r]]:ptr[docs[r] + 1]]
        qt[r, 4:] = rng.choice(d, 4)
    npost = ix.query_postings(qt)
    qd = torch.from_numpy(qt).to(dev)
    ts = timeit(lambda: ix.search(qd, K), warm=1, it=3)
    ms = min(ts)
    print(f"bm25 1M docs Q=4096 k=100: {ts} ms -> {Q/ms*1e3:.0f} q/s; postings/query={npost/Q:.0f}; {npost*4/ms/1e6:.1f} GB/s algorithmic (4B/posting)", flush=True)
