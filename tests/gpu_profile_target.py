"""Profiling target: one dense search (4096 q, 1M x 768 bf16) and one BM25 search at BASELINE size.
Usage: python tests/gpu_profile_target.py [bm25_queries]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ano_rag_b200
from ano_rag_b200 import engine, bm25_build as bb, synth

dev = torch.device("cuda:0")
nbq = int(sys.argv[1]) if len(sys.argv) > 1 else 296
N, D, Q, K, V = 1_000_000, 768, 4096, 100, 50_000
dense = engine.DenseIndex(D, storage="bf16", normalize=False)
for _, rows in synth.dense_rows(0, N, D, dev):
    dense.add(rows)
q = synth.dense_queries(Q, D, dev)
ptr, tok = synth.bm25_corpus(N, V)
qt = synth.bm25_queries(Q, ptr, tok, V)
arr = bb.build_arrays(ptr, tok, V)
bm = engine.BM25Index(N, arr.term_ptr, arr.post_doc, arr.post_gcode, arr.stats.idf, arr.gtab)
qd = torch.from_numpy(qt[:nbq]).to(dev)
for _ in range(2):
    dense.search(q, K, query_planes=1)
    bm.search(qd, K)
torch.cuda.synchronize()
print("done", flush=True)
