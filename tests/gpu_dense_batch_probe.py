"""Development probe: dense stage time at several batch sizes (cfg3 corpus), checked against a torch matmul top-k."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ano_rag_b200
from ano_rag_b200 import engine, synth, _native as nat
dev = torch.device("cuda:0")
N, D, K = 1_000_000, 768, 100
ix = engine.DenseIndex(D, storage="bf16", normalize=False)
chunks = []
for _, rows in synth.dense_rows(0, N, D, dev):
    ix.add(rows); chunks.append(rows)
X = torch.cat(chunks)
for Q in ([int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else (4096, 2048, 1024, 512, 128)):
    q = synth.dense_queries(Q, D, dev, n_rows=N)
    qq = q[:64]
    ref = (qq.float() @ X.float().T).topk(K, dim=1)
    base = None
    for cross in ("1", "0"):
        os.environ["ARB_DENSE_CROSS"] = cross
        for _ in range(3): s, i = ix.search(q, K, query_planes=1)
        torch.cuda.synchronize(); nat.profile(True)
        for _ in range(10): ix.search(q, K, query_planes=1)
        torch.cuda.synchronize(); prof = nat.profile_collect(); nat.profile(False)
        ms = prof["dense_topk"][0] / prof["dense_topk"][1]
        agree = np.mean([len(set(a.tolist()) & set(b.tolist())) / K for a, b in zip(ref.indices.cpu().numpy(), i[:64].cpu().numpy())])
        same = base is None or (torch.equal(base[0], s) and torch.equal(base[1], i))
        base = base or (s, i)
        print(f"Q={Q} cross-list bound={cross}: {ms:.3f} ms -> {2*Q*N*D/ms/1e9:.0f} TFLOP/s  agreement with torch fp32 matmul "
              f"{agree:.4f}  identical to bound=0: {same}", flush=True)
