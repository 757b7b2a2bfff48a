"""Small end-to-end pass over every kernel family, sized for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck --error-exitcode 1 python tests/gpu_sanitize.py
Results are also checked against the per-query BM25 kernel / a torch matmul so a silent failure is visible."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ano_rag_b200  # noqa: F401
from ano_rag_b200 import engine, synth
from ano_rag_b200.utils.bm25_search import GpuBM25
from ano_rag_b200.retrieval.hybrid_search import HybridSearcher
from ano_rag_b200.pipeline import HybridPipeline
from ano_rag_b200.query import HybridRescorer, EnhancedHybridSearchV2, KhopReranker

which = set((sys.argv[1] if len(sys.argv) > 1 else "dense,ivf,bm25,fuse,rescore").split(","))
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
N, D, Q, K = 3000, 96, 130, 10

if "dense" in which:
    for storage, planes in (("bf16", 1), ("fp32", 3)):
        x = rng.standard_normal((N, D)).astype(np.float32)
        q = rng.standard_normal((Q, D)).astype(np.float32)
        ix = engine.DenseIndex(D, storage=storage, normalize=True)
        ix.add(x)
        s, i = ix.search(torch.from_numpy(q).to(dev), K, query_planes=planes)
        xn = x / np.linalg.norm(x, axis=1, keepdims=True); qn = q / np.linalg.norm(q, axis=1, keepdims=True)
        top = np.argsort(-(qn @ xn.T), axis=1)[:, :K]
        agree = np.mean([len(set(a) & set(b)) / K for a, b in zip(top, i.cpu().numpy())])
        print(f"dense {storage}: agreement with numpy {agree:.3f}", flush=True)
        assert agree > (0.99 if storage == "fp32" else 0.9)
        cand = rng.integers(-1, N, (Q, 7))
        ix.score_rows(q, cand)
        ix.close()

if "ivf" in which:
    x = rng.standard_normal((N, D)).astype(np.float32)
    ix = engine.DenseIndex(D, storage="bf16", normalize=True)
    ix.train(x, 16, niter=3)
    ix.add(x)
    q = torch.from_numpy(rng.standard_normal((Q, D)).astype(np.float32)).to(dev)
    a = ix.ivf_search(q, K, 16)[1]
    b = ix.search(q, K)[1]
    print("ivf all lists == flat:", bool((a == b).all()), flush=True)
    ix.ivf_search(q, K, 3)
    ix.close()

if "bm25" in which:
    n, v = 20_000, 800
    ptr, tok = synth.bm25_corpus(n, v, 20, seed=1)
    qt = synth.bm25_queries(Q, ptr, tok, v, seed=2)
    qt[5, :] = 0
    g = GpuBM25.from_token_ids(ptr, tok, v, variant="okapi")
    os.environ["ARB_BM25_MODE"] = "query"
    s0, i0 = g.search_terms(qt, 20)
    os.environ["ARB_BM25_MODE"] = "batch"
    s1, i1 = g.search_terms(qt, 20)
    print("bm25 batch == per-query:", np.array_equal(s0, s1) and np.array_equal(i0, i1), flush=True)
    assert np.array_equal(s0, s1) and np.array_equal(i0, i1)
    g.index.scores(qt[:4])
    g.close()

if "fuse" in which:
    cfg = {"retrieval": {"candidate_pool": K, "hybrid": {"fusion_method": "rrf", "rrf_k": 60, "weights": {"dense": 1.0, "bm25": 1.0}}}}
    n, v = 5000, 300
    ptr, tok = synth.bm25_corpus(n, v, 15, seed=3)
    g = GpuBM25.from_token_ids(ptr, tok, v, variant="okapi")
    ix = engine.DenseIndex(D, storage="bf16", normalize=False)
    ix.add(rng.standard_normal((n, D)).astype(np.float32))
    for method in ("rrf", "linear"):
        cfg["retrieval"]["hybrid"]["fusion_method"] = method
        pipe = HybridPipeline(ix, g.index, HybridSearcher(cfg), k_dense=K, k_bm25=K, pool=K, device=dev)
        q = torch.from_numpy(rng.standard_normal((Q, D)).astype(np.float32)).to(dev)
        qt = torch.from_numpy(synth.bm25_queries(Q, ptr, tok, v, seed=4)).to(dev)
        ids, fin, src, cnt = pipe.search_device(q, qt)
        print(f"fuse {method}: mean pool fill {cnt.float().mean().item():.1f}", flush=True)
    sc = torch.rand((3, Q, K), device=dev).sort(dim=2, descending=True)[0]
    idl = torch.randint(0, 10_000, (3, Q, K), device=dev)
    engine.merge_topk(sc, idl)
    engine.merge_topk(sc.double(), idl)
    g.close(); ix.close()

if "rescore" in which:
    words = [f"t{i}" for i in range(12)]
    def cands(n):
        return [{"note_id": f"n{i}", "content": " ".join(rng.choice(words, 4)), "hop_no": int(rng.integers(1, 4))} for i in range(n)]
    batch = [{"candidates": cands(n), "vector_scores": rng.random(n).tolist(), "bm25_scores": rng.random(n).tolist(),
              "must_have_terms": ["t1"], "boost_entities": ["t2"], "boost_predicates": ["t3"]} for n in (1, 40, 300)]
    for m in ("linear", "rrf"):
        HybridRescorer(m).rescore_batch(batch)
    EnhancedHybridSearchV2().search_batch([dict(b, dense_scores=b["vector_scores"], sparse_scores=b["bm25_scores"]) for b in batch])
    KhopReranker().rerank_batch([dict(b, dense_scores=b["vector_scores"], focused_scores=b["vector_scores"], path_entities=["t1"]) for b in batch])
    print("rescore stages ran", flush=True)
torch.cuda.synchronize()
print("SANITIZE PASS DONE", flush=True)
