"""Profiling target for the batched BM25 pipeline: cfg3 corpus, [nq] queries, two searches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ano_rag_b200
from ano_rag_b200 import engine, bm25_build as bb, synth

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N, K, V = 1_000_000, 100, 50_000
ptr, tok = synth.bm25_corpus(N, V)
qt = synth.bm25_queries(4096, ptr, tok, V)
arr = bb.build_arrays(ptr, tok, V)
bm = engine.BM25Index(N, arr.term_ptr, arr.post_doc, arr.post_gcode, arr.stats.idf, arr.gtab)
qd = torch.from_numpy(qt[:nq]).to("cuda:0")
for _ in range(2):
    bm.search(qd, K)
torch.cuda.synchronize()
print("done", flush=True)
