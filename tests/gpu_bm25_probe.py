"""Development probe (not the bench): the cfg3 BM25 stage alone, per-stage device times.

    python tests/gpu_bm25_probe.py [--notes N] [--queries Q] [--iters I]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ano_rag_b200  # noqa: F401
from ano_rag_b200 import _native as nat, bm25_build as bb, engine, synth

ap = argparse.ArgumentParser()
ap.add_argument("--notes", type=int, default=1_000_000)
ap.add_argument("--vocab", type=int, default=50_000)
ap.add_argument("--queries", type=int, default=4096)
ap.add_argument("--topk", type=int, default=100)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--check", type=int, default=0, help="compare this many queries with the per-query kernel")
a = ap.parse_args()
dev = torch.device("cuda:0")
t0 = time.time()
ptr, tok = synth.bm25_corpus(a.notes, a.vocab)
qt = synth.bm25_queries(a.queries, ptr, tok, a.vocab)
arr = bb.build_arrays(ptr, tok, a.vocab, variant="okapi")
t1 = time.time()
ix = engine.BM25Index(a.notes, arr.term_ptr, arr.post_doc, arr.post_gcode, arr.stats.idf, arr.gtab)
torch.cuda.synchronize()
print(f"host arrays {t1 - t0:.1f}s, device index {time.time() - t1:.2f}s, postings {len(arr.post_doc)}", flush=True)
qd = torch.from_numpy(qt).to(dev)
ix.search(qd, a.topk)
torch.cuda.synchronize()
nat.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    s, i = ix.search(qd, a.topk)
e1.record()
torch.cuda.synchronize()
prof = nat.profile_collect()
nat.profile(False)
ms = e0.elapsed_time(e1) / a.iters
print(f"bm25 stage {ms:.3f} ms/batch ({a.queries} queries) ->", {k: round(v[0] / max(1, v[1]), 3) for k, v in prof.items() if v[1]}, flush=True)
if os.environ.get("ARB_BM25_STATS"):
    print("batch stats:", nat.bm25_batch_stats(), flush=True)
if a.check:
    os.environ["ARB_BM25_MODE"] = "query"
    s0, i0 = ix.search(qd[: a.check].contiguous(), a.topk)
    ok = bool(torch.equal(s0, s[: a.check]) and torch.equal(i0, i[: a.check]))
    print(f"check vs per-query kernel on {a.check} queries: {'identical' if ok else 'MISMATCH'}", flush=True)
