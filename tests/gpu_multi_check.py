"""Multi-GPU parity (run under torchrun on N GPUs): the row-sharded pipeline must give the same fused
top-k as one unsharded index (identical scores, ties broken by global id).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/gpu_multi_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import ano_rag_b200
from ano_rag_b200 import bm25_build as bb, engine, sharded, synth
from ano_rag_b200.pipeline import HybridPipeline
from ano_rag_b200.retrieval.hybrid_search import HybridSearcher

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, D, Q, K, V = 100_000, 256, 512, 100, 5000
cfg = {"retrieval": {"candidate_pool": K, "hybrid": {"fusion_method": "rrf", "rrf_k": 60, "weights": {"dense": 1.0, "bm25": 1.0}}}}
ptr, tok = synth.bm25_corpus(N, V, mean_len=30)
qt = torch.from_numpy(synth.bm25_queries(Q, ptr, tok, V)).to(dev)
q = synth.dense_queries(Q, D, dev, n_rows=N)
ok = True
for method in ("rrf", "linear"):
    cfg["retrieval"]["hybrid"]["fusion_method"] = method
    lo, hi = sharded.shard_range(N, world, rank)
    dense = engine.DenseIndex(D, storage="bf16", normalize=False)
    for _, rows in synth.dense_rows(lo, hi - lo, D, dev):
        dense.add(rows)
    sptr, stok = ptr[lo:hi + 1] - ptr[lo], tok[ptr[lo]:ptr[hi]]
    stats = sharded.global_bm25_stats(sptr, stok, V, sharded.token_bases(len(stok), device=dev), device=dev)
    arr = bb.build_arrays(sptr, stok, V, stats=stats)
    bm = engine.BM25Index(hi - lo, arr.term_ptr, arr.post_doc, arr.post_gcode, stats.idf, arr.gtab)
    pipe = HybridPipeline(dense, bm, HybridSearcher(cfg), k_dense=K, k_bm25=K, pool=K, id_base=lo, query_planes=1, device=dev)
    res = {}
    for ex in ("peer", "nccl"):                             # NVLink peer-load merge vs NCCL all-to-all + merge
        pipe.exchange = ex
        for _ in range(3):
            res[ex] = pipe.search_device(q, qt)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            res[ex] = pipe.search_device(q, qt)
        e1.record(); torch.cuda.synchronize()
        if rank == 0:
            print(f"[{method}] exchange={ex}: {e0.elapsed_time(e1) / 20:.3f} ms/batch", flush=True)
    same_ex = all(bool((a == b).all().item()) for a, b in zip(res["peer"][:2], res["nccl"][:2]))
    same_ex &= bool((res["peer"][3] == res["nccl"][3]).all().item())
    flag = torch.tensor([int(same_ex)], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"[{method}] peer exchange == nccl exchange on every rank: {bool(flag.item())}", flush=True)
    ok &= bool(flag.item())
    ids, fin, src, cnt = res["peer"]
    qlo, qhi = pipe.query_slice(Q)
    # gather every rank's slice on rank 0
    per = (Q + world - 1) // world
    pad_ids = torch.full((per, K), -7, dtype=torch.int64, device=dev); pad_ids[: qhi - qlo] = ids
    pad_fin = torch.zeros((per, K), dtype=torch.float64, device=dev); pad_fin[: qhi - qlo] = fin
    all_ids = torch.empty((world, per, K), dtype=torch.int64, device=dev)
    all_fin = torch.empty((world, per, K), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(all_ids, pad_ids)
    dist.all_gather_into_tensor(all_fin, pad_fin)
    if rank == 0:
        full_dense = engine.DenseIndex(D, storage="bf16", normalize=False)
        for _, rows in synth.dense_rows(0, N, D, dev):
            full_dense.add(rows)
        farr = bb.build_arrays(ptr, tok, V)
        assert np.array_equal(farr.stats.idf, stats.idf) and farr.stats.avgdl == stats.avgdl
        full_bm = engine.BM25Index(N, farr.term_ptr, farr.post_doc, farr.post_gcode, farr.stats.idf, farr.gtab)
        ref = HybridPipeline(full_dense, full_bm, HybridSearcher(cfg), k_dense=K, k_bm25=K, pool=K, query_planes=1, device=dev)
        ref.world, ref.rank = 1, 0                          # unsharded reference on this rank only
        r_ids, r_fin, _, _ = ref.search_device(q, qt)
        got_ids = all_ids.reshape(world * per, K)[:Q]
        got_fin = all_fin.reshape(world * per, K)[:Q]
        same_ids = bool((got_ids == r_ids).all().item())
        same_fin = bool((got_fin == r_fin).all().item())
        print(f"[{method}] world={world}: fused ids identical={same_ids} scores bit-identical={same_fin}", flush=True)
        ok &= same_ids and same_fin
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MULTI-GPU PARITY", "OK" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)
