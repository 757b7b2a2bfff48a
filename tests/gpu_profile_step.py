"""Profiling target: two full cfg3 steps (dense top-k -> BM25 top-k -> fusion) at the bench's exact shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ano_rag_b200
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
for _ in range(2):
    pipe.search_device(q, qt)
torch.cuda.synchronize()
print("done", flush=True)
