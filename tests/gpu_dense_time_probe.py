"""Development probe: dense cfg3 timing for {single, pair} x {pre-pass off, on}; results must be identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ano_rag_b200
from ano_rag_b200 import engine, synth, _native as nat
dev = torch.device("cuda:0")
N, D, Q, K = 1_000_000, 768, 4096, 100
ix = engine.DenseIndex(D, storage="bf16", normalize=False)
for _, rows in synth.dense_rows(0, N, D, dev):
    ix.add(rows)
q = synth.dense_queries(Q, D, dev, n_rows=N)
ref = None
for pair in (0, 1):
    for pre in (0, 1):
        os.environ["ARB_DENSE_PAIR"] = str(pair); os.environ["ARB_DENSE_PREPASS"] = str(pre)
        for _ in range(3): s, i = ix.search(q, K, query_planes=1)
        torch.cuda.synchronize(); nat.profile(True)
        for _ in range(10): ix.search(q, K, query_planes=1)
        torch.cuda.synchronize(); prof = nat.profile_collect(); nat.profile(False)
        ms = prof["dense_topk"][0] / prof["dense_topk"][1]
        s, i = s.cpu().numpy(), i.cpu().numpy()
        if ref is None: ref = (s, i)
        same = np.array_equal(ref[0], s) and np.array_equal(ref[1], i)
        print(f"pair={pair} prepass={pre}: {ms:.3f} ms -> {2*Q*N*D/ms/1e9:.0f} TFLOP/s  identical={same}", flush=True)
