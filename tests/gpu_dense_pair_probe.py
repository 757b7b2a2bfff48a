"""Development probe: CTA-pair dense kernel vs the single-CTA kernel (bit-identical results expected) + timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ano_rag_b200
from ano_rag_b200 import engine, synth, _native as nat

dev = torch.device("cuda:0")
def run(ix, q, k, pair, planes=None):
    os.environ["ARB_DENSE_PAIR"] = "1" if pair else "0"
    s, i = ix.search(q, k, query_planes=planes)
    torch.cuda.synchronize()
    return s.cpu().numpy(), i.cpu().numpy()

ok = True
for (n, d, nq, k, storage) in [(50_000, 128, 256, 10, "bf16"), (100_003, 768, 1000, 100, "bf16"), (30_000, 96, 300, 200, "fp32"), (70_000, 200, 513, 20, "fp32")]:
    g = torch.Generator(device=dev); g.manual_seed(n)
    x = torch.nn.functional.normalize(torch.randn((n, d), generator=g, device=dev), dim=1)
    q = torch.nn.functional.normalize(x[torch.randint(0, n, (nq,), generator=g, device=dev)] + 0.3 * torch.randn((nq, d), generator=g, device=dev) / d ** 0.5, dim=1)
    ix = engine.DenseIndex(d, storage=storage, normalize=False)
    ix.add(x.to(torch.bfloat16) if storage == "bf16" else x)
    qq = q.to(torch.bfloat16) if storage == "bf16" else q
    s0, i0 = run(ix, qq, k, False)
    s1, i1 = run(ix, qq, k, True)
    same = np.array_equal(i0, i1) and np.array_equal(s0, s1)
    print(f"n={n} d={d} nq={nq} k={k} {storage}: pair == single: {same}", flush=True)
    if not same:
        bad = np.flatnonzero((i0 != i1).any(1) | (s0 != s1).any(1))
        print("  differing queries:", len(bad), bad[:8], i0[bad[0]][:6], i1[bad[0]][:6], s0[bad[0]][:4], s1[bad[0]][:4])
    ok &= same
    ix.close()
if len(sys.argv) > 1 and sys.argv[1] == "time":
    N, D, Q, K = 1_000_000, 768, 4096, 100
    ix = engine.DenseIndex(D, storage="bf16", normalize=False)
    for _, rows in synth.dense_rows(0, N, D, dev):
        ix.add(rows)
    q = synth.dense_queries(Q, D, dev, n_rows=N)
    for pair in (False, True):
        os.environ["ARB_DENSE_PAIR"] = "1" if pair else "0"
        for _ in range(3): ix.search(q, K, query_planes=1)
        torch.cuda.synchronize()
        nat.profile(True)
        for _ in range(10): ix.search(q, K, query_planes=1)
        torch.cuda.synchronize()
        prof = nat.profile_collect(); nat.profile(False)
        ms = prof["dense_topk"][0] / prof["dense_topk"][1]
        print(f"pair={pair}: dense_topk {ms:.3f} ms -> {2*Q*N*D/ms/1e9:.0f} TFLOP/s", flush=True)
print("PAIR PROBE", "OK" if ok else "FAIL")
sys.exit(0 if ok else 1)
