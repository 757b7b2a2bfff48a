"""Development probe at BASELINE cfg5 size: 10M x 768 bf16 IVF-4096 (topic-cluster centroids), nprobe 32,
4096 queries, top-100: time and recall@100 against the flat kernel.  usage: gpu_ivf_probe.py [n] [nlist] [nprobe]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ano_rag_b200
from ano_rag_b200 import engine, _native as nat

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
NLIST = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
NPROBE = int(sys.argv[3]) if len(sys.argv) > 3 else 32
D, Q, K = 768, 4096, 100
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1234)
centres = torch.nn.functional.normalize(torch.randn((NLIST, D), generator=g, device=dev), dim=1)
ix = engine.DenseIndex(D, storage="bf16", normalize=False)
sums = torch.zeros((NLIST, D), device=dev); cnts = torch.zeros(NLIST, device=dev)
t0 = time.time()
CH = 500_000
qsrc = None
for s in range(0, N, CH):
    m = min(CH, N - s)
    topic = torch.randint(0, NLIST, (m,), generator=g, device=dev)
    x = centres[topic] + 0.3 * torch.randn((m, D), generator=g, device=dev) / D ** 0.5
    x = torch.nn.functional.normalize(x, dim=1)
    sums.index_add_(0, topic, x); cnts.index_add_(0, topic, torch.ones(m, device=dev))
    if qsrc is None:
        qsrc = x[:Q].clone()
    ix.add(x.to(torch.bfloat16))
torch.cuda.synchronize()
print(f"corpus {N} x {D} built in {time.time()-t0:.1f}s", flush=True)
cent = sums / cnts.clamp(min=1)[:, None]                       # per-cluster mean (doc/clustering.py:194)
t0 = time.time()
ix.set_centroids(cent.contiguous())
torch.cuda.synchronize()
print(f"centroids installed + rows quantised in {time.time()-t0:.2f}s", flush=True)
q = torch.nn.functional.normalize(qsrc + 0.2 * torch.randn((Q, D), generator=g, device=dev) / D ** 0.5, dim=1).to(torch.bfloat16)

def timeit(fn, warm=2, it=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it, out

t0 = time.time()
ix.ivf_search(q[:128], K, NPROBE); torch.cuda.synchronize()
print(f"first search (regroup) {time.time()-t0:.2f}s", flush=True)
nat.profile(True)
ms_ivf, (s_ivf, i_ivf) = timeit(lambda: ix.ivf_search(q, K, NPROBE))
prof = nat.profile_collect(); nat.profile(False)
print({k: round(v[0] / max(1, v[1]), 3) for k, v in prof.items() if v[1]})
ms_flat, (s_flat, i_flat) = timeit(lambda: ix.search(q, K), warm=1, it=2)
a, b = i_ivf.cpu().numpy(), i_flat.cpu().numpy()
rec = np.mean([len(set(x) & set(y)) / K for x, y in zip(a, b)])
print(f"IVF-{NLIST} nprobe={NPROBE}: {ms_ivf:.2f} ms/batch ({Q/ms_ivf*1e3:.0f} q/s); flat: {ms_flat:.2f} ms/batch; recall@{K} vs flat = {rec:.4f}", flush=True)
