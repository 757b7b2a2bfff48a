"""Development probe: hybrid step latency with the dense and BM25 sources on one stream vs two (cfg3 shapes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ano_rag_b200  # noqa: F401
from ano_rag_b200 import engine, bm25_build as bb, synth
from ano_rag_b200.pipeline import HybridPipeline
from ano_rag_b200.retrieval.hybrid_search import HybridSearcher

dev = torch.device("cuda:0")
N, D, Q, K, V = 1_000_000, 768, 4096, 100, 50_000
dense = engine.DenseIndex(D, storage="bf16", normalize=False)
for _, rows in synth.dense_rows(0, N, D, dev):
    dense.add(rows)
q = synth.dense_queries(Q, D, dev, n_rows=N)
ptr, tok = synth.bm25_corpus(N, V)
qt = torch.from_numpy(synth.bm25_queries(Q, ptr, tok, V)).to(dev)
arr = bb.build_arrays(ptr, tok, V)
bm = engine.BM25Index(N, arr.term_ptr, arr.post_doc, arr.post_gcode, arr.stats.idf, arr.gtab)
pipe = HybridPipeline(dense, bm, HybridSearcher({"retrieval": {"candidate_pool": K, "hybrid": {
    "fusion_method": "rrf", "rrf_k": 60, "weights": {"dense": 1.0, "bm25": 1.0}}}}), k_dense=K, k_bm25=K, pool=K, query_planes=1, device=dev)

def med(fn, reps=20):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2], out

for nq in (1, 8, 128, 512, 1024, 2048, 4096):
    qe, qq = q[:nq].contiguous(), qt[:nq].contiguous()
    pipe.overlap_max_queries = 0
    t0, r0 = med(lambda: pipe.search_device(qe, qq))
    pipe.overlap_max_queries = 1 << 30
    t1, r1 = med(lambda: pipe.search_device(qe, qq))
    same = all(torch.equal(a, b) for a, b in zip(r0[:2], r1[:2]))
    print(f"nq={nq:5d}: one stream {t0:.3f} ms, two streams {t1:.3f} ms, identical={same}", flush=True)
