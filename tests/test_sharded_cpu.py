"""World-size-2 gloo test of the sharding host logic (runs on CPU): global BM25 statistics from
per-rank shards must equal the single-corpus statistics bit for bit, and the shard index arrays
must hold exactly the single index's postings."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ano_rag_b200  # noqa: F401
from ano_rag_b200 import bm25_build as bb
from ano_rag_b200 import sharded, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_docs, vocab, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ptr, tok = synth.bm25_corpus(n_docs, vocab, mean_len=12, seed=5)
        lo, hi = sharded.shard_range(n_docs, world, rank)
        sptr = ptr[lo:hi + 1] - ptr[lo]
        stok = tok[ptr[lo]:ptr[hi]]
        base = sharded.token_bases(len(stok))
        assert base == int(ptr[lo])
        stats = sharded.global_bm25_stats(sptr, stok, vocab, base)
        arr = bb.build_arrays(sptr, stok, vocab, stats=stats)
        np.savez(os.path.join(out, f"r{rank}.npz"), idf=stats.idf, df=stats.df, avgdl=stats.avgdl,
                 average_idf=stats.average_idf, n=stats.n_docs, term_ptr=arr.term_ptr, post_doc=arr.post_doc + lo,
                 g=arr.gtab[arr.post_gcode])
    finally:
        dist.destroy_process_group()


def test_global_stats_and_shards_world2(tmp_path):
    n_docs, vocab, world = 700, 90, 2
    mp.spawn(_worker, args=(world, _free_port(), n_docs, vocab, str(tmp_path)), nprocs=world, join=True)
    ptr, tok = synth.bm25_corpus(n_docs, vocab, mean_len=12, seed=5)
    whole = bb.build_arrays(ptr, tok, vocab)
    shards = [np.load(os.path.join(str(tmp_path), f"r{r}.npz")) for r in range(world)]
    for z in shards:
        assert int(z["n"]) == n_docs and float(z["avgdl"]) == whole.stats.avgdl
        assert float(z["average_idf"]) == whole.stats.average_idf
        assert np.array_equal(z["idf"], whole.stats.idf) and np.array_equal(z["df"], whole.stats.df)
    for t in range(vocab):
        docs = np.concatenate([z["post_doc"][z["term_ptr"][t]:z["term_ptr"][t + 1]] for z in shards])
        g = np.concatenate([z["g"][z["term_ptr"][t]:z["term_ptr"][t + 1]] for z in shards])
        s, e = whole.term_ptr[t], whole.term_ptr[t + 1]
        assert np.array_equal(docs, whole.post_doc[s:e])
        assert np.array_equal(g, whole.gtab[whole.post_gcode[s:e]])


def test_shard_range_covers_rows():
    for n, w in [(10, 3), (1_000_000, 8), (5, 8), (0, 2)]:
        spans = [sharded.shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _exchange_worker(rank, world, port, nq, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ano_rag_b200.pipeline import HybridPipeline
        pipe = HybridPipeline.__new__(HybridPipeline)            # only the exchange logic: no device objects
        pipe.world, pipe.rank, pipe.group = world, rank, None
        local = torch.arange(nq * k, dtype=torch.int64).reshape(nq, k) + 1_000_000 * rank    # this rank's [nq, k] list
        got = pipe._exchange(local, -1)
        np.save(os.path.join(out, f"x{rank}.npy"), got.numpy())
        np.save(os.path.join(out, f"s{rank}.npy"), np.array(pipe.query_slice(nq)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nq", [10, 7])
def test_query_slice_exchange_world2(tmp_path, nq):
    """all-to-all of query slices: rank r receives, from every rank, the rows of ITS slice (padded with -1)."""
    world, k = 2, 3
    mp.spawn(_exchange_worker, args=(world, _free_port(), nq, k, str(tmp_path)), nprocs=world, join=True)
    per = (nq + world - 1) // world
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"x{r}.npy"))
        lo, hi = np.load(os.path.join(str(tmp_path), f"s{r}.npy"))
        assert got.shape == (world, per, k) and (lo, hi) == (min(nq, r * per), min(nq, (r + 1) * per))
        for src in range(world):
            want = np.arange(nq * k).reshape(nq, k)[lo:hi] + 1_000_000 * src
            assert np.array_equal(got[src, : hi - lo], want)
            assert (got[src, hi - lo:] == -1).all()


def test_grid_layout_covers_every_row_and_query_exactly_once():
    """N GPUs = R row shards x N/R query groups (bench.py layouts): inside a group the shards tile the corpus, across groups the
    query slices tile the batch, and the ranks' output slices tile the batch too (every query is finished by exactly one rank)."""
    for world in (1, 2, 4, 8):
        for r_shards in [r for r in (1, 2, 4, 8) if world % r == 0 and r <= world]:
            for n_rows, n_q in ((1_000_000, 4096), (1001, 37), (5, 3)):
                rows = np.zeros((world // r_shards, n_rows), dtype=np.int32)
                served = np.zeros(n_q, dtype=np.int32)
                out = np.zeros(n_q, dtype=np.int32)
                for rank in range(world):
                    g = sharded.grid_layout(world, rank, r_shards, n_rows, n_q)
                    assert g["peers"] == list(range(g["group"] * r_shards, (g["group"] + 1) * r_shards)) and rank in g["peers"]
                    rows[g["group"], g["rows"][0]:g["rows"][1]] += 1
                    if g["shard"] == 0:
                        served[g["queries"][0]:g["queries"][1]] += 1
                    out[g["out_queries"][0]:g["out_queries"][1]] += 1
                    assert g["queries"][0] <= g["out_queries"][0] <= g["out_queries"][1] <= g["queries"][1]
                assert np.all(rows == 1) and np.all(served == 1) and np.all(out == 1), (world, r_shards, n_rows, n_q)
    with pytest.raises(ValueError):
        sharded.grid_layout(8, 0, 3, 10, 10)
