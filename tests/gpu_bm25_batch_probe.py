"""Development probe: batched BM25 pipeline vs the per-query kernel at cfg3 size (parity on every query,
sub-kernel timing, fallback / emission statistics).  usage: gpu_bm25_batch_probe.py [n_docs] [nq]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ARB_BM25_STATS"] = "1"
import numpy as np
import torch
import ano_rag_b200
from ano_rag_b200 import engine, bm25_build as bb, synth, _native as nat

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
V, K = 50_000, 100
dev = torch.device("cuda:0")
t0 = time.time()
ptr, tok = synth.bm25_corpus(N, V, 40)
qt = synth.bm25_queries(Q, ptr, tok, V)
arr = bb.build_arrays(ptr, tok, V, variant="okapi")
print(f"host build {time.time()-t0:.1f}s postings={len(arr.post_doc)} n_g={len(arr.gtab)}", flush=True)
t0 = time.time()
ix = engine.BM25Index(N, arr.term_ptr, arr.post_doc, arr.post_gcode, arr.stats.idf, arr.gtab)
torch.cuda.synchronize()
print(f"device index {time.time()-t0:.1f}s", flush=True)
qd = torch.from_numpy(qt).to(dev)

def run(mode, it=3):
    os.environ["ARB_BM25_MODE"] = mode
    s, i = ix.search(qd, K)
    torch.cuda.synchronize()
    nat.bm25_batch_stats(reset=True)
    nat.profile(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(it):
        s, i = ix.search(qd, K)
    ev1.record()
    torch.cuda.synchronize()
    prof = nat.profile_collect()
    nat.profile(False)
    print(mode, f"{ev0.elapsed_time(ev1)/it:.3f} ms/batch", {k: round(v[0] / it, 3) for k, v in prof.items() if v[1]},
          nat.bm25_batch_stats(), flush=True)
    return s.cpu().numpy(), i.cpu().numpy()

sq, iq = run("query")
sb, ib = run("batch")
same_i = (iq == ib).all(axis=1)
same_s = (sq == sb).all(axis=1)
print(f"batch == per-query: ids {same_i.sum()}/{Q} scores {same_s.sum()}/{Q}", flush=True)
if not (same_i.all() and same_s.all()):
    bad = np.flatnonzero(~(same_i & same_s))[:5]
    for r in bad:
        j = np.flatnonzero((iq[r] != ib[r]) | (sq[r] != sb[r]))[:4]
        print("query", r, "terms", qt[r], "first diffs at", j, "query-mode", iq[r][j], sq[r][j], "batch", ib[r][j], sb[r][j])
    sys.exit(1)
