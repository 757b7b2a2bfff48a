"""First-contact diagnostics on a B200: every kernel against the oracle, verbose.
Run under gpurun:  python tests/gpu_probe.py [stage ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ano_rag_b200
from ano_rag_b200 import engine
from ano_rag_b200 import bm25_build as bb
from oracle import dense as od, bm25 as obm, fusion as ofu

stages = sys.argv[1:] or ["dense", "bm25", "fuse"]
torch.cuda.init()
print("device", torch.cuda.get_device_name(0), flush=True)


def dense_case(n, d, nq, k, storage, qp, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = (x[rng.integers(0, n, nq)] + 0.5 * rng.standard_normal((nq, d))).astype(np.float32)
    ix = engine.DenseIndex(d, storage=storage, normalize=True)
    ix.add(x)
    t0 = time.time()
    s, i = ix.search(q, k, query_planes=qp)
    dt = time.time() - t0
    xs, qs = od.preprocess_vectors(x), od.preprocess_vectors(q)
    if storage == "bf16":
        xs = od.bf16_round(xs)
    if qp == 1:
        qs = od.bf16_round(qs)
    ws, wi = od.flat_ip_topk(qs, xs, k)
    ok, nm, tot, worst = od.compare_topk(i, s, wi, ws, tol=1e-5)
    print(f"dense n={n} d={d} nq={nq} k={k} {storage} qp={qp}: ok={ok} id-match={nm}/{tot} worst|ds|={worst:.3e} t={dt*1e3:.1f}ms", flush=True)
    if not ok:
        bad = np.nonzero((i != wi).any(axis=1))[0][:3]
        for r in bad:
            print(" row", r, "got", i[r][:8], s[r][:8], "\n     want", wi[r][:8], ws[r][:8])
    return ok


if "dense" in stages:
    allok = True
    allok &= dense_case(1000, 64, 5, 10, "bf16", 1)
    allok &= dense_case(1000, 64, 5, 10, "fp32", 3)
    allok &= dense_case(10000, 384, 200, 20, "fp32", 3)
    allok &= dense_case(10000, 384, 200, 20, "bf16", 1)
    allok &= dense_case(70001, 100, 300, 100, "bf16", 1)
    allok &= dense_case(300, 768, 130, 256, "fp32", 3)
    allok &= dense_case(50, 32, 3, 100, "bf16", 1)
    print("DENSE", "OK" if allok else "FAIL", flush=True)

if "bm25" in stages:
    rng = np.random.default_rng(1)
    allok = True
    for (n, v, variant) in [(500, 60, "okapi"), (40000, 3000, "okapi"), (40000, 3000, "simple")]:
        lens = np.maximum(1, rng.poisson(20, n))
        ptr = np.zeros(n + 1, np.int64); ptr[1:] = np.cumsum(lens)
        p = 1.0 / np.arange(1, v + 1); p /= p.sum()
        tok = rng.choice(v, size=int(ptr[-1]), p=p)
        arr = bb.build_arrays(ptr, tok, v, variant=variant)
        ix = engine.BM25Index(n, arr.term_ptr, arr.post_doc, arr.post_gcode, arr.stats.idf, arr.gtab)
        ora = obm.ArrayBM25Oracle(ptr, tok, variant=variant)
        nq, mq, k = 16, 8, 50
        qt = rng.choice(v, size=(nq, mq), p=p).astype(np.int32)
        qt[0, 3:] = -1; qt[1, 2] = qt[1, 0]
        full = ix.scores(qt)
        s, i = ix.search(qt, k)
        for r in range(nq):
            want = ora.get_scores([t for t in qt[r] if t >= 0])
            if not np.array_equal(full[r], want):
                allok = False
                bad = np.nonzero(full[r] != want)[0]
                print(f" bm25 {variant} n={n} q{r}: {len(bad)} score mismatches, first doc {bad[0]} got {full[r][bad[0]]!r} want {want[bad[0]]!r}")
            ws, wi = obm.topk_positive(want, k)
            if not (np.array_equal(i[r], wi) and np.array_equal(s[r], ws)):
                allok = False
                print(f" bm25 {variant} n={n} q{r}: topk mismatch got {i[r][:6]} {s[r][:3]} want {wi[:6]} {ws[:3]}")
        nrm = ix.scores(qt, n_out=n + 5, normalise=True)
        want0 = ora.get_scores([t for t in qt[2] if t >= 0]); want0 = np.concatenate([want0, np.zeros(5)])
        if want0.max() > 0: want0 = want0 / want0.max()
        if not np.array_equal(nrm[2], want0):
            allok = False; print(" bm25 normalise mismatch")
        print(f"bm25 {variant} n={n} v={v}: done ok={allok}", flush=True)
    print("BM25", "OK" if allok else "FAIL", flush=True)

if "fuse" in stages:
    import random
    rng = random.Random(3)
    allok = True
    for method in ("linear", "rrf"):
        nq, kd, kb, pool = 7, 20, 30, 25
        w = {"dense": 1.0, "bm25": 0.5, "graph": 0.35, "path": 0.1}
        d_ids = np.full((nq, kd), -1, np.int64); d_sc = np.zeros((nq, kd), np.float32)
        b_ids = np.full((nq, kb), -1, np.int64); b_sc = np.zeros((nq, kb), np.float64)
        for r in range(nq):
            nd = rng.randint(0, kd); nb = rng.randint(0, kb)
            d_ids[r, :nd] = rng.sample(range(60), nd); d_sc[r, :nd] = np.sort(np.random.rand(nd).astype(np.float32))[::-1]
            b_ids[r, :nb] = rng.sample(range(60), nb); b_sc[r, :nb] = np.sort(np.random.rand(nb) * 9)[::-1]
        oi, of, osrc, oc = engine.fuse_batch(method, w, 60, pool, dense=(d_ids, d_sc), bm25=(b_ids, b_sc))
        cfg = {"retrieval": {"candidate_pool": pool, "hybrid": {"fusion_method": method, "rrf_k": 60, "weights": w}}}
        for r in range(nq):
            dl = [(int(a), float(b)) for a, b in zip(d_ids[r], d_sc[r]) if a >= 0]
            bl = [(int(a), float(b)) for a, b in zip(b_ids[r], b_sc[r]) if a >= 0]
            want = ofu.HybridSearcherOracle(cfg).fuse(dl, bl)
            got = [(int(oi[r, j]), float(of[r, j])) for j in range(oc[r])]
            if got != [(x["note_id"], x["final_similarity"]) for x in want]:
                allok = False
                print(f" fuse {method} q{r} mismatch\n  got  {got[:5]}\n  want {[(x['note_id'], x['final_similarity']) for x in want][:5]}")
        print(f"fuse {method}: ok={allok}", flush=True)
    print("FUSE", "OK" if allok else "FAIL", flush=True)
