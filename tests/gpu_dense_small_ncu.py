"""Development probe for an ncu launch list: dense search at small batch sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ano_rag_b200
from ano_rag_b200 import engine, synth
dev = torch.device("cuda:0")
N, D, K = 1_000_000, 768, 100
ix = engine.DenseIndex(D, storage="bf16", normalize=False)
for _, rows in synth.dense_rows(0, N, D, dev):
    ix.add(rows)
for Q in (4096, 512, 128):
    q = synth.dense_queries(Q, D, dev, n_rows=N)
    for _ in range(3):
        ix.search(q, K, query_planes=1)
    torch.cuda.synchronize()
print("done")
