"""Development probe: latency of the dense search (1 M x 768 bf16, k = 100) at small batch sizes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ano_rag_b200  # noqa: F401
from ano_rag_b200 import engine, synth

dev = torch.device("cuda:0")
N, D, K = int(os.environ.get("N", 1_000_000)), 768, 100
ix = engine.DenseIndex(D, storage="bf16", normalize=False)
for _, rows in synth.dense_rows(0, N, D, dev):
    ix.add(rows)
q_all = synth.dense_queries(4096, D, dev, n_rows=N)
for nq in (1, 8, 16, 64, 128, 512, 4096):
    q = q_all[:nq].contiguous()
    for _ in range(5):
        ix.search(q, K, query_planes=1)
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ix.search(q, K, query_planes=1); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"nq={nq:5d}: median {ts[len(ts)//2]:.3f} ms  min {ts[0]:.3f} ms", flush=True)
