#!/usr/bin/env python
"""Headline benchmark: hybrid queries/sec @ top-100 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (N=1): BASELINE.json configs[2], which contains configs[1] as its dense
stage -- 1M synthetic atomic notes x 768-d bf16 (exact IP top-100) + BM25 over 1M
docs (Zipf vocab 50k, ~40 tokens) + RRF(k=60) fusion to the top-100, 4096-query
batch.  One "step" = one batch through dense top-k -> BM25 top-k -> fusion.
N>1 (strong scaling, same corpus and batch): N GPUs = R row shards x N/R query groups.  The
default picks the fewest row shards that fit one GPU's HBM -- for this 1M-note corpus R = 1: every
GPU holds the corpus and serves 1/N of the batch, nothing to exchange -- and ALSO measures the
fully row-sharded layout (R = N: local top-k per source, candidates exchanged over NVLink by the
merge kernel's peer loads, merge, fuse), reported in the same line under "row_sharded".
`--row-shards R` forces a layout; `--workload cfg4` is the corpus that only fits row-sharded.

`value`  = queries/s with the batch already resident in HBM (device-timed, max over ranks).
`e2e`    = same metric through the public API with HOST buffers (H2D of the batch and D2H
           of the fused result inside the timed region).
`--impl reference` times the reference's CPU algorithm (the oracle port: faiss/rank_bm25
are not installable offline) on the host cores, on a bounded query sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hybrid_queries_per_sec_top100"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4"],
                    help="cfg3 (default, the driver's line): hybrid over 1M notes, row-sharded when N>1 (strong scaling). "
                         "cfg4: dense only, 12.5M x 1024-d bf16 notes PER GPU, 8192 queries (N=8 is BASELINE configs[3])")
    ap.add_argument("--notes", type=int, default=None)
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--queries", type=int, default=None)
    ap.add_argument("--topk", type=int, default=100)
    ap.add_argument("--vocab", type=int, default=50_000)
    ap.add_argument("--row-shards", default="auto",
                    help="N GPUs = R row shards x N/R query groups.  'auto' (default): the fewest row shards whose share of "
                         "the corpus fits one GPU's HBM budget (cfg3: 1 -> the corpus is replicated and the batch is split; "
                         "cfg4: N); an integer dividing N forces R (R = N is the fully row-sharded layout)")
    ap.add_argument("--hbm-budget-gb", type=float, default=140.0, help="per-GPU budget for index data when choosing R")
    ap.add_argument("--cpu-sample", type=int, default=64, help="queries per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def resolve_workload(a):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.workload == "cfg4":
        a.notes = a.notes or 12_500_000 * world
        a.dim = a.dim or 1024
        a.queries = a.queries or 8192
    else:
        a.notes = a.notes or 1_000_000
        a.dim = a.dim or 768
        a.queries = a.queries or 4096
    return a


def choose_row_shards(a, world):
    """R row shards x (world / R) query groups."""
    if a.row_shards != "auto":
        r = int(a.row_shards)
        if r < 1 or world % r:
            raise SystemExit(f"--row-shards {r} must divide the number of GPUs ({world})")
        return r
    if a.workload == "cfg4":
        return world
    corpus_gb = (a.notes * a.dim * 2 + a.notes * 40 * 12) / 1e9          # bf16 rows + postings, planes and build scratch
    for r in range(1, world + 1):
        if world % r == 0 and corpus_gb / r <= a.hbm_budget_gb:
            return r
    return world


def workload_config(a, n_gpus, row_shards=None):
    row_shards = row_shards or n_gpus
    groups = n_gpus // row_shards
    layout = "none" if n_gpus == 1 else f"{row_shards} row shard(s) x {groups} query group(s)"
    if a.workload == "cfg4":
        return {"workload": f"cfg4 dense: {a.notes} notes x {a.dim}-d bf16 row-sharded over {n_gpus} GPU(s) "
                            f"({a.notes // n_gpus} rows each), exact IP local top-{a.topk} + all-gather merge, {a.queries}-query batch",
                "notes": a.notes, "dim": a.dim, "queries": a.queries, "top_k": a.topk,
                "sharding": layout,
                "l2": "inputs larger than L2 (corpus shard 25.6 GB bf16 vs 126 MB L2)"}
    return {
        "workload": f"cfg3 hybrid: {a.notes} notes x {a.dim}-d bf16 flat-IP top-{a.topk} + BM25 (Zipf vocab {a.vocab}, "
                    f"~40 tok/doc, okapi fp64) + RRF k=60 -> fused top-{a.topk}, {a.queries}-query batch",
        "notes": a.notes, "dim": a.dim, "queries": a.queries, "top_k": a.topk, "vocab": a.vocab,
        "fusion": "rrf", "rrf_k": 60, "weights": {"dense": 1.0, "bm25": 1.0},
        "sharding": layout,
        "l2": "inputs larger than L2 (corpus 1.5 GB bf16 + 138 MB postings vs 126 MB L2)",
    }


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "tflops": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json; sustained bf16 figure: kernel timed inside a long step)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# --------------------------------------------------------------------------------------
class CpuPath:
    def __init__(self, a, dense_rows_f32=None, queries_f32=None, doc_ptr=None, doc_tok=None, q_terms=None):
        import numpy as np
        import torch
        from oracle import bm25 as obm, fusion as ofu
        from ano_rag_b200 import synth          # generators only (no kernels)
        self.np, self.a = np, a
        if dense_rows_f32 is None:
            g = torch.Generator().manual_seed(synth.SEED_CORPUS)
            x = torch.nn.functional.normalize(torch.randn((a.notes, a.dim), generator=g), dim=1)
            dense_rows_f32 = x.to(torch.bfloat16).to(torch.float32).numpy()
            j = torch.randint(0, min(a.notes, 65536), (a.queries,), generator=g)
            q = x[j] + 0.5 * torch.randn((a.queries, a.dim), generator=g) / (a.dim ** 0.5)
            queries_f32 = torch.nn.functional.normalize(q, dim=1).to(torch.bfloat16).to(torch.float32).numpy()
        if doc_ptr is None:
            doc_ptr, doc_tok = synth.bm25_corpus(a.notes, a.vocab)
            q_terms = synth.bm25_queries(a.queries, doc_ptr, doc_tok, a.vocab)
        self.x, self.q, self.q_terms = dense_rows_f32, queries_f32, q_terms
        self.bm = obm.ArrayBM25Oracle(doc_ptr, doc_tok, variant="okapi")
        self.fuser = ofu.HybridSearcherOracle({"retrieval": {"candidate_pool": a.topk, "hybrid": {
            "fusion_method": "rrf", "rrf_k": 60, "weights": {"dense": 1.0, "bm25": 1.0}}}})
        self.obm = obm

    def step(self, lo, n):
        """dense flat-IP top-k + BM25 top-k + fuse for queries [lo, lo+n) -- the reference's CPU algorithm."""
        from oracle import dense as od
        np = self.np
        k = self.a.topk
        sc, ids = od.flat_ip_topk(self.q[lo:lo + n], self.x, k)
        out = []
        for r in range(n):
            full = self.bm.get_scores([t for t in self.q_terms[lo + r] if t >= 0])
            bs, bi = self.obm.topk_positive(full, k)
            dense = [(int(i), float(s)) for i, s in zip(ids[r], sc[r]) if i >= 0]
            bm25 = [(int(i), float(s)) for i, s in zip(bi, bs) if i >= 0]
            out.append(self.fuser.fuse(dense, bm25))
        return out


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    cpu = CpuPath(a)
    n = min(a.cpu_sample, a.queries)
    nblk = max(1, a.queries // n)
    for w in range(a.warmup):
        cpu.step((w % nblk) * n, n)
    t0 = time.perf_counter()
    for s in range(a.steps):
        cpu.step((s % nblk) * n, n)
    dt = time.perf_counter() - t0
    qps = a.steps * n / dt
    sample = f"{n} of {a.queries} queries per step against the full {a.notes}-note corpus (dense blocked sgemm + argpartition " \
             f"on all cores via BLAS; BM25 CSR numpy port of rank_bm25 arithmetic, single thread like the library; HybridSearcher.fuse port)"
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic", "config": workload_config(a, a.gpus),
            "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "restated reference CPU path (faiss / rank_bm25 unavailable offline); a step is a bounded sample"}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200 import _native as nat, bm25_build as bb, engine, synth
    from ano_rag_b200.pipeline import HybridPipeline
    from ano_rag_b200.retrieval.hybrid_search import HybridSearcher

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly one JSON line: anything libraries print meanwhile (NCCL's version banner ...) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    hybrid = a.workload == "cfg3"
    q_emb_all = synth.dense_queries(a.queries, a.dim, dev, n_rows=a.notes)          # bf16, resident
    doc_ptr = doc_tok = q_terms_all = full = None
    if hybrid:
        doc_ptr, doc_tok = synth.bm25_corpus(a.notes, a.vocab)
        q_terms_all = synth.bm25_queries(a.queries, doc_ptr, doc_tok, a.vocab)
        full = bb.build_arrays(doc_ptr, doc_tok, a.vocab, variant="okapi")
    searcher = HybridSearcher({"retrieval": {"candidate_pool": a.topk, "hybrid": {
        "fusion_method": "rrf", "rrf_k": 60, "weights": {"dense": 1.0, "bm25": 1.0}}}})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def build_layout(n_shards):
        """world = n_shards row shards x n_groups query groups: rank r holds row shard r % n_shards and serves the
        query slice of group r // n_shards; the ranks of one group exchange and merge their candidates."""
        n_groups = world // n_shards
        shard, grp = rank % n_shards, rank // n_shards
        per = (a.notes + n_shards - 1) // n_shards
        row0, row1 = min(a.notes, shard * per), min(a.notes, (shard + 1) * per)
        qper = (a.queries + n_groups - 1) // n_groups
        q0, q1 = min(a.queries, grp * qper), min(a.queries, (grp + 1) * qper)
        dense = engine.DenseIndex(a.dim, storage="bf16", normalize=False)
        for _, rows in synth.dense_rows(row0, row1 - row0, a.dim, dev):
            dense.add(rows)
        bm25 = None
        if hybrid:
            if n_shards > 1:
                sptr = doc_ptr[row0:row1 + 1] - doc_ptr[row0]
                arr = bb.build_arrays(sptr, doc_tok[doc_ptr[row0]:doc_ptr[row1]], a.vocab, stats=full.stats)
            else:
                arr = full
            bm25 = engine.BM25Index(row1 - row0, arr.term_ptr, arr.post_doc, arr.post_gcode, arr.stats.idf, arr.gtab)
        sub = None
        if world > 1 and n_shards > 1:
            subs = [dist.new_group(list(range(g * n_shards, (g + 1) * n_shards))) for g in range(n_groups)]
            sub = subs[grp]
        pipe = HybridPipeline(dense, bm25, searcher, k_dense=a.topk, k_bm25=a.topk, pool=a.topk, id_base=row0,
                              query_planes=1, group=sub, device=dev)
        if n_shards == 1:
            pipe.world, pipe.rank = 1, 0                                            # replicated corpus: nothing to exchange
        q_terms_h = q_terms_all[q0:q1] if hybrid else None
        return {"pipe": pipe, "dense": dense, "bm25": bm25, "n_shards": n_shards, "n_groups": n_groups, "rows": (row0, row1),
                "queries": (q0, q1), "q_emb": q_emb_all[q0:q1].contiguous(), "q_terms_h": q_terms_h,
                "q_terms": torch.from_numpy(np.ascontiguousarray(q_terms_h)).to(dev) if hybrid else None,
                "local_postings": bm25.query_postings(q_terms_h) if hybrid and q1 > q0 else 0}

    def measure(lay, sampler=None):
        pipe, q_emb, q_terms = lay["pipe"], lay["q_emb"], lay["q_terms"]
        n_warm = max(3, a.warmup)
        for _ in range(n_warm):
            pipe.search_device(q_emb, q_terms)
        barrier()
        # a fresh box keeps paging driver / library code in for the first steps (launch gaps of several ms): keep warming,
        # untimed, until two consecutive steps agree within 10 % (at most 12 extra steps); the count is reported
        prev = None
        for _ in range(12):
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record()
            pipe.search_device(q_emb, q_terms)
            w1.record()
            barrier()
            n_warm += 1
            cur = max_over_ranks(w0.elapsed_time(w1))
            if prev is not None and abs(cur - prev) <= 0.1 * min(cur, prev):
                break
            prev = cur
        nat.launch_count(reset=True)
        nat.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            pipe.search_device(q_emb, q_terms)
        e1.record()
        barrier()
        launches = nat.launch_count(reset=True)
        prof = nat.profile_collect()
        nat.profile(False)
        ms_step = max_over_ranks(e0.elapsed_time(e1)) / a.steps
        clocks = sampler.stop() if sampler is not None else None
        # ---- end to end: pinned host inputs, host outputs
        q_emb_h = q_emb.to(torch.float32).cpu().pin_memory()                          # encoder output is fp32 on the host
        q_terms_p = torch.from_numpy(np.ascontiguousarray(lay["q_terms_h"])).pin_memory() if hybrid else None
        for _ in range(3):
            pipe.search_host(q_emb_h, q_terms_p)
        barrier()
        e0.record()
        for _ in range(a.steps):
            res = pipe.search_host(q_emb_h, q_terms_p)
        e1.record()
        barrier()
        e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / a.steps
        h2d = q_emb_h.numel() * 4 + (q_terms_p.numel() * 4 if hybrid else 0)
        d2h = sum(int(r.nbytes) for r in res)
        return {"ms_step": ms_step, "value": a.queries / (ms_step * 1e-3), "warmup": n_warm, "launches": int(launches),
                "prof": prof, "clocks": clocks, "res": res,
                "e2e": {"value": a.queries / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms}}

    # ---- the headline layout (clock sampler runs from the warm-up on: nvidia-smi needs ~0.5 s to start)
    n_shards = choose_row_shards(a, world)
    lay = build_layout(n_shards)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    m = measure(lay, sampler if rank == 0 else None)
    pipe, prof, res, e2e, clocks = lay["pipe"], m["prof"], m["res"], m["e2e"], m["clocks"]
    ms_step, value, n_warm, launches = m["ms_step"], m["value"], m["warmup"], m["launches"]
    row0, row1 = lay["rows"]
    q_local = lay["queries"][1] - lay["queries"][0]
    local_postings, q_terms_h, q_emb = lay["local_postings"], q_terms_all, q_emb_all
    lo, hi = pipe.query_slice(q_local)
    lo, hi = lo + lay["queries"][0], hi + lay["queries"][0]

    # ---- the fully row-sharded layout beside it when the headline replicates the corpus (SURVEY 8(e) exchange + merge)
    row_sharded = None
    if world > 1 and n_shards != world and a.row_shards == "auto":
        lay["dense"].close()
        if lay["bm25"] is not None:
            lay["bm25"].close()
        lay2 = build_layout(world)
        m2 = measure(lay2)
        row_sharded = {"row_shards": world, "query_groups": 1, "value": m2["value"], "unit": UNIT, "ms_per_step": m2["ms_step"],
                       "e2e": m2["e2e"], "gpu_launches": m2["launches"], "exchange": lay2["pipe"].exchange}

    # ---- roofline of the dominant kernel (timed live, CUDA events on the launch stream)
    peaks = load_peaks()
    d_ms = prof["dense_topk"][0] / max(1, prof["dense_topk"][1])
    b_ms = prof["bm25_topk"][0] / max(1, prof["bm25_topk"][1])
    n_local = row1 - row0
    dense_tflops = 2.0 * q_local * n_local * a.dim / (d_ms * 1e-3) / 1e12 if d_ms > 0 else 0.0
    # bytes/query = sum over query tokens of df(t) * (4 B doc id + 4 B impact) + k * 8 B of output.
    # (The index itself stores 4 B per posting and the batched pipeline re-uses term planes across the
    # queries of a batch from shared memory, so DRAM traffic is far below this figure: see `traffic`.)
    bm25_bytes = local_postings * 8 + q_local * a.topk * 8
    bm25_gbs = bm25_bytes / (b_ms * 1e-3) / 1e9 if b_ms > 0 else 0.0

    def per_launch(kind):
        return prof[kind][0] / max(1, prof[kind][1]) if kind in prof else 0.0
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
    stages = {
        "dense_topk": {"ms": d_ms, "bound": "tensor", "achieved": dense_tflops, "peak": peaks["tflops"], "unit": "TFLOP/s",
                       "frac": dense_tflops / peaks["tflops"], "algorithmic_flops": 2.0 * q_local * n_local * a.dim},
        "bm25_topk": {"ms": b_ms, "bound": "hbm", "achieved": bm25_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                      "frac": bm25_gbs / peaks["hbm_gbs"], "algorithmic_bytes": bm25_bytes,
                      "postings_per_query": local_postings / max(1, q_local),
                      "kernels_ms": {"sparse_pass": per_launch("bm25_sparse"), "plane_pass": per_launch("bm25_planes"),
                                     "final_merge_and_fallback": per_launch("bm25_final")}},
        "fuse": {"ms": prof["fuse"][0] / max(1, prof["fuse"][1])},
        "dense_merge": {"ms": prof["dense_merge"][0] / max(1, prof["dense_merge"][1])},
    }
    dom = "bm25_topk" if b_ms >= d_ms else "dense_topk"
    roofline = {"kernel": dom, "bound": stages[dom]["bound"], "achieved": stages[dom]["achieved"],
                "peak": stages[dom]["peak"], "unit": stages[dom]["unit"], "frac": stages[dom]["frac"],
                "traffic": traffic.get(dom) if world == 1 else None, "peak_source": peaks["source"], "launch_ms": stages[dom]["ms"]}

    # ---- CPU baseline beside it (rank 0, N=1, bounded sample)
    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline and hybrid:
        xs = np.concatenate([r.to(torch.float32).cpu().numpy() for _, r in synth.dense_rows(0, a.notes, a.dim, dev)])
        cpu = CpuPath(a, xs, q_emb.to(torch.float32).cpu().numpy(), doc_ptr, doc_tok, q_terms_h)
        n = min(a.cpu_sample, a.queries)
        cpu.step(0, min(8, n))                                                     # warm BLAS threads
        t0 = time.perf_counter()
        ref = cpu.step(0, n)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": n / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"{n} of {a.queries} queries against the full corpus, {dt:.1f} s (oracle port: blocked sgemm + "
                                  f"argpartition on all cores, rank_bm25 arithmetic over CSR in numpy, HybridSearcher.fuse port)"}
        # the same sample doubles as an end-to-end parity check of the step
        ids = res[0]
        agree = 0
        for r in range(min(n, hi - lo)):
            want = [e["note_id"] for e in ref[lo + r]] if lo == 0 else []
            agree += len(set(want) & set(ids[r][:len(want)].tolist())) / max(1, len(want))
        cpu_baseline["fused_top100_agreement_with_oracle"] = agree / max(1, min(n, hi - lo))

    if rank == 0:
        line = {"metric": METRIC if hybrid else "dense_queries_per_sec_top100", "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": n_warm,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if hybrid else "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": workload_config(a, n_gpus, n_shards), "e2e": e2e,
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "stages": stages,
                "cpu_baseline": cpu_baseline, "row_sharded": row_sharded, "impl": "b200",
                "dtype_note": "dense: bf16 operands, fp32 accumulate on tcgen05; BM25 and fusion: fp64"}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    a = resolve_workload(parse_args())
    if a.impl == "reference":
        return run_reference(a)
    return run_b200(a)


if __name__ == "__main__":
    sys.exit(main())
