"""BASELINE.json configs[4] as a measured workload (``bench.py --workload cfg5``): 10 M notes x 768-d bf16 in an
IVF-4096 topic-cluster index (centroids = per-topic means of the rows, the reference's ``doc/clustering.py:194``
handed to ``IndexIVFFlat``, vector_store/vector_index.py:84-91), nprobe 32, top-100, recall against the exact flat scan
of the same rows, fused (RRF k=60) with BM25 over 10 M docs.  One GPU: 15.4 GB of rows + their list-major copy +
~13 GB of BM25 index fit one B200's HBM.

A step = one 4096-query batch through IVF top-100 -> BM25 top-100 -> fusion.  Parity at this size is checked through
properties the size does not change: the fused lists equal the oracle's fusion of the two per-source lists (sampled
queries, ordered (id, score) pairs exact), the batched BM25 pipeline equals the one-CTA-per-query kernel (another code
path, pinned to the oracle at the sizes the oracle finishes) on a sample, nprobe = nlist equals the flat scan, and every
list is ordered.  With ``--notes`` <= 2 M the BM25 lists are also compared with the CPU oracle itself.
"""
from __future__ import annotations

import os
import time

METRIC = "hybrid_queries_per_sec_top100"
UNIT = "queries/s"


class _IvfAsDense:
    """``DenseIndex`` seen through its IVF search: what ``HybridPipeline`` calls for the dense source."""

    def __init__(self, ix, nprobe):
        self.ix, self.nprobe = ix, nprobe

    def search(self, q, k, query_planes=None, id_base=0, out=None):
        s, i = self.ix.ivf_search(q, k, self.nprobe, query_planes=query_planes, id_base=id_base)
        if out is not None:
            out[0].copy_(s)
            out[1].copy_(i)
            return out
        return s, i


def run_cfg5(a, dev, peaks, sampler):
    import numpy as np
    import torch
    from ano_rag_b200 import _native as nat, bm25_build as bb, engine, synth
    from ano_rag_b200.pipeline import HybridPipeline
    from ano_rag_b200.retrieval.hybrid_search import HybridSearcher
    from oracle import fusion as ofu

    n, d, nq, k = a.notes, a.dim, a.queries, a.topk
    nlist, nprobe = 4096, 32
    t_setup = time.time()
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    centres = torch.nn.functional.normalize(torch.randn((nlist, d), generator=g, device=dev), dim=1)
    ix = engine.DenseIndex(d, storage="bf16", normalize=False)
    sums = torch.zeros((nlist, d), device=dev)
    cnts = torch.zeros(nlist, device=dev)
    qsrc = None
    for s in range(0, n, 500_000):
        m = min(500_000, n - s)
        topic = torch.randint(0, nlist, (m,), generator=g, device=dev)
        x = torch.nn.functional.normalize(centres[topic] + 0.3 * torch.randn((m, d), generator=g, device=dev) / d ** 0.5, dim=1)
        sums.index_add_(0, topic, x)
        cnts.index_add_(0, topic, torch.ones(m, device=dev))
        if qsrc is None:
            qsrc = x[:nq].clone()
        ix.add(x.to(torch.bfloat16))
        del x, topic
    ix.set_centroids((sums / cnts.clamp(min=1)[:, None]).contiguous())                    # the cluster mean, doc/clustering.py:194
    q = torch.nn.functional.normalize(qsrc + 0.2 * torch.randn((nq, d), generator=g, device=dev) / d ** 0.5, dim=1).to(torch.bfloat16)
    del sums, cnts, qsrc
    ix.ivf_search(q[:128].contiguous(), k, nprobe, query_planes=1)                        # first search regroups the rows list-major
    torch.cuda.synchronize()
    dense_s = time.time() - t_setup

    t0 = time.time()
    doc_ptr, doc_tok = synth.bm25_corpus(n, a.vocab)
    q_terms_h = synth.bm25_queries(nq, doc_ptr, doc_tok, a.vocab)
    arr = bb.build_arrays(doc_ptr, doc_tok, a.vocab, variant="okapi")
    bm = engine.BM25Index(n, arr.term_ptr, arr.post_doc, arr.post_gcode, arr.stats.idf, arr.gtab)
    postings = bm.query_postings(q_terms_h)
    q_terms = torch.from_numpy(np.ascontiguousarray(q_terms_h)).to(dev)
    bm25_s = time.time() - t0
    del arr

    cfg = {"retrieval": {"candidate_pool": k, "hybrid": {"fusion_method": "rrf", "rrf_k": 60, "weights": {"dense": 1.0, "bm25": 1.0}}}}
    pipe = HybridPipeline(_IvfAsDense(ix, nprobe), bm, HybridSearcher(cfg), k_dense=k, k_bm25=k, pool=k, query_planes=1, device=dev)

    def ev_time(fn, steps, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, out

    warm = max(3, a.warmup)
    sampler.start()
    for _ in range(warm):
        pipe.search_device(q, q_terms)
    torch.cuda.synchronize()
    nat.launch_count(reset=True)
    ms_step, res = ev_time(lambda: pipe.search_device(q, q_terms), a.steps, 0)
    launches = nat.launch_count(reset=True)
    clocks = sampler.stop()
    # stage durations from a pass with both sources on ONE stream (the headline overlaps them on two)
    pipe.overlap_max_queries = 0
    nat.profile(True)
    ms_one_stream, _ = ev_time(lambda: pipe.search_device(q, q_terms), max(2, a.steps // 2), 1)
    prof = nat.profile_collect()
    nat.profile(False)
    ms_two_streams = ms_step
    if ms_one_stream < ms_step:                                                            # headline = the faster stream mode
        ms_step = ms_one_stream
    else:
        pipe.overlap_max_queries = 1 << 30
    stages = {kk: v[0] / max(1, v[1]) for kk, v in prof.items() if v[1]}
    q_h = q.to(torch.float32).cpu().pin_memory()
    qt_p = torch.from_numpy(np.ascontiguousarray(q_terms_h)).pin_memory()
    e2e_ms, res_h = ev_time(lambda: pipe.search_host(q_h, qt_p), a.steps, 3)
    h2d = q_h.numel() * 4 + qt_p.numel() * 4
    d2h = sum(int(r.nbytes) for r in res_h)

    # ---- per-stage numbers and recall
    ivf_ms, (vs, vi) = ev_time(lambda: ix.ivf_search(q, k, nprobe, query_planes=1), max(3, a.steps // 2), 1)
    bm_ms, (bs, bi) = ev_time(lambda: bm.search(q_terms, k), max(3, a.steps // 2), 1)
    n_rec = min(512, nq)
    flat_ms, (fs, fi) = ev_time(lambda: ix.search(q[:n_rec].contiguous(), k, query_planes=1), 1, 1)
    vi_h, fi_h = vi.cpu().numpy(), fi.cpu().numpy()
    recall = float(np.mean([len(set(x) & set(y)) / k for x, y in zip(vi_h[:n_rec], fi_h)]))
    full_ms, (es, ei) = ev_time(lambda: ix.ivf_search(q[:64].contiguous(), k, nlist, query_planes=1), 1, 0)
    ei_h, es_h = ei.cpu().numpy(), es.cpu().numpy()
    fs_h = fs.cpu().numpy()
    all_lists_equal_flat = bool(all(set(x) == set(y) or np.allclose(np.sort(sx), np.sort(sy), atol=1e-5)
                                    for x, y, sx, sy in zip(ei_h, fi_h[:64], es_h, fs_h[:64])))

    # ---- parity through size-independent properties
    ids, fin, src, cnt = (t.cpu().numpy() for t in res)
    vs_h, bs_h, bi_h = vs.cpu().numpy(), bs.cpu().numpy(), bi.cpu().numpy()
    fuser = ofu.HybridSearcherOracle(cfg)
    rng = np.random.default_rng(15)
    sample = np.sort(rng.choice(nq, min(64, nq), replace=False))
    fused_ok = 0
    for r in sample:
        want = fuser.fuse([(int(i), float(s)) for i, s in zip(vi_h[r], vs_h[r]) if i >= 0],
                          [(int(i), float(s)) for i, s in zip(bi_h[r], bs_h[r]) if i >= 0])
        got = [(int(ids[r, j]), float(fin[r, j])) for j in range(int(cnt[r]))]
        fused_ok += got == [(e["note_id"], e["final_similarity"]) for e in want]
    os.environ["ARB_BM25_MODE"] = "query"
    try:
        s0, i0 = bm.search(q_terms[:64].contiguous(), k)
    finally:
        del os.environ["ARB_BM25_MODE"]
    bm25_paths_equal = bool(np.array_equal(i0.cpu().numpy(), bi_h[:64]) and np.array_equal(s0.cpu().numpy(), bs_h[:64]))
    ordered = bool(np.all(np.diff(vs_h, axis=1) <= 0) and np.all(np.diff(bs_h, axis=1) <= 0))
    bm25_oracle = None
    if n <= 2_000_000:
        from oracle import bm25 as obm
        ora = obm.ArrayBM25Oracle(doc_ptr, doc_tok, variant="okapi")
        bm25_oracle = 0
        for r in sample[:16]:
            ws, wi = obm.topk_positive(ora.get_scores([t for t in q_terms_h[r] if t >= 0]), k)
            bm25_oracle += bool(np.array_equal(bi_h[r], wi) and np.array_equal(bs_h[r], ws))
        bm25_oracle = {"queries": 16, "bit_exact": int(bm25_oracle)}
    parity = {"fused_equals_oracle_fusion_of_the_source_lists": f"{fused_ok}/{len(sample)}",
              "bm25_batched_pipeline_equals_per_query_kernel_on_64_queries": bm25_paths_equal,
              "ivf_nprobe_eq_nlist_equals_flat_on_64_queries": all_lists_equal_flat, "lists_ordered": ordered,
              "bm25_vs_cpu_oracle": bm25_oracle,
              "green": bool(fused_ok == len(sample) and bm25_paths_equal and all_lists_equal_flat and ordered
                            and (bm25_oracle is None or bm25_oracle["bit_exact"] == 16))}

    # ---- rooflines: the IVF scan reads nprobe/nlist of the rows once per query TILE, the BM25 stage its postings
    scanned_rows = nq * nprobe / nlist * n                                                   # rows scored, summed over queries
    ivf_tflops = 2.0 * scanned_rows * d / (ivf_ms * 1e-3) / 1e12
    bm25_bytes = postings * 8 + nq * k * 8
    bm25_gbs = bm25_bytes / (bm_ms * 1e-3) / 1e9
    dom_is_bm25 = bm_ms >= ivf_ms
    roofline = ({"kernel": "bm25 stage (member + screening passes)", "bound": "hbm", "achieved": bm25_gbs, "peak": peaks["hbm_gbs"],
                 "unit": "GB/s", "frac": bm25_gbs / peaks["hbm_gbs"], "traffic": None, "launch_ms": bm_ms,
                 "note": "ALGORITHMIC bytes (8 B per posting a query touches); plane tiles are re-used across the batch"}
                if dom_is_bm25 else
                {"kernel": "dense_topk_kernel<IVF> (tcgen05, list-major rows)", "bound": "tensor", "achieved": ivf_tflops,
                 "peak": peaks["tflops_burst"], "unit": "TFLOP/s", "frac": ivf_tflops / peaks["tflops_burst"], "traffic": None,
                 "launch_ms": ivf_ms, "note": "useful flops only (rows of the probed lists x real queries)"})
    line = {"metric": METRIC, "value": nq / (ms_step * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": warm,
            "ms_per_step": ms_step, "ms_per_step_one_stream": ms_one_stream, "ms_per_step_two_streams": ms_two_streams, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"cfg5 hybrid: {n} notes x {d}-d bf16 IVF-{nlist} (topic-mixture rows, centroids = per-topic means), "
                                   f"nprobe {nprobe}, top-{k} + BM25 over {n} docs (Zipf vocab {a.vocab}, okapi fp64) + RRF k=60, {nq}-query batch",
                       "notes": n, "dim": d, "queries": nq, "top_k": k, "nlist": nlist, "nprobe": nprobe, "vocab": a.vocab,
                       "l2": "inputs larger than L2 (15 GB of rows, ~13 GB of BM25 index vs 126 MB L2)"},
            "e2e": {"value": nq / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "stages": {"ivf_topk_ms": ivf_ms, "bm25_topk_ms": bm_ms, "profile_ms": stages,
                       "ivf_useful_tflops": ivf_tflops, "bm25_algorithmic_gbs": bm25_gbs, "bm25_postings_per_query": postings / nq},
            "recall_at_100_vs_flat": {"value": recall, "queries": n_rec, "flat_ms_for_those_queries": flat_ms},
            "parity": parity, "cpu_baseline": None,
            "setup_s": {"dense_rows_centroids_regroup": dense_s, "bm25_host_arrays_and_device_index": bm25_s},
            "impl": "b200"}
    ix.close()
    bm.close()
    return line
