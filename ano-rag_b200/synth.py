"""Seeded synthetic inputs for the BASELINE.json configurations.

Dense rows are generated in fixed chunks of ``CHUNK`` rows from a generator keyed
by (seed, chunk index), so any sharding of the corpus sees identical rows.  With
``host_rng=True`` the generator runs on the CPU whatever the target device, so the
B200 arm and the CPU reference arm of ``bench.py`` score bit-identical inputs (cfg3);
the 100 M-row cfg4 shards are generated on the device.  This module imports nothing
of the package: ``bench.py --impl reference`` loads it by file path without mapping
the CUDA library.  BM25 documents are Poisson(mean_len)-long token streams drawn from
a Zipf(s=1) law over ``vocab`` term ids; queries mix 4 Zipf-drawn tokens with 4
tokens of a random document so that real matches exist.
"""
from __future__ import annotations

import numpy as np
import torch

CHUNK = 65536
SEED_CORPUS, SEED_QUERIES, SEED_TEXT = 1234, 1235, 1236


def _chunk(seed: int, c: int, d: int, device, host_rng: bool = False) -> torch.Tensor:
    gen_dev = "cpu" if host_rng else device
    g = torch.Generator(device=gen_dev)
    g.manual_seed(seed * 1_000_003 + c)
    x = torch.randn((CHUNK, d), generator=g, device=gen_dev, dtype=torch.float32)
    return torch.nn.functional.normalize(x, dim=1)


def dense_rows(row0: int, n: int, d: int, device, dtype=torch.bfloat16, seed: int = SEED_CORPUS, host_rng: bool = False):
    """Yield (first_row, tensor) pieces covering rows [row0, row0+n): unit-norm rows rounded to ``dtype``."""
    r = row0
    end = row0 + n
    while r < end:
        c = r // CHUNK
        x = _chunk(seed, c, d, device, host_rng)
        lo = r - c * CHUNK
        hi = min(CHUNK, end - c * CHUNK)
        yield r, x[lo:hi].to(dtype).to(device)
        r = c * CHUNK + hi


def dense_queries(nq: int, d: int, device, dtype=torch.bfloat16, n_rows: int = CHUNK, seed: int = SEED_QUERIES,
                  corpus_seed: int = SEED_CORPUS, host_rng: bool = False) -> torch.Tensor:
    """normalise(X[j] + 0.5*N(0,1)) for random rows j of the first corpus chunk."""
    gen_dev = "cpu" if host_rng else device
    g = torch.Generator(device=gen_dev)
    g.manual_seed(seed)
    base = _chunk(corpus_seed, 0, d, device, host_rng)[: min(CHUNK, n_rows)]
    j = torch.randint(0, base.shape[0], (nq,), generator=g, device=gen_dev)
    q = base[j] + 0.5 * torch.randn((nq, d), generator=g, device=gen_dev) / (d ** 0.5)
    return torch.nn.functional.normalize(q, dim=1).to(dtype).to(device)


def zipf_cdf(vocab: int) -> np.ndarray:
    p = 1.0 / np.arange(1, vocab + 1, dtype=np.float64)
    return np.cumsum(p / p.sum())


def bm25_corpus(n_docs: int, vocab: int = 50_000, mean_len: int = 40, seed: int = SEED_TEXT):
    """-> (doc_ptr int64 [n+1], doc_tok int64 [total]) token-id corpus."""
    rng = np.random.default_rng(seed)
    lens = np.maximum(1, rng.poisson(mean_len, n_docs))
    ptr = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    tok = np.searchsorted(zipf_cdf(vocab), rng.random(int(ptr[-1]))).astype(np.int64)
    np.clip(tok, 0, vocab - 1, out=tok)
    return ptr, tok


def bm25_queries(nq: int, doc_ptr: np.ndarray, doc_tok: np.ndarray, vocab: int = 50_000, seed: int = SEED_QUERIES):
    """[nq, 8] int32: 4 Zipf tokens + 4 tokens of a random document."""
    rng = np.random.default_rng(seed)
    q = np.empty((nq, 8), dtype=np.int32)
    q[:, :4] = np.clip(np.searchsorted(zipf_cdf(vocab), rng.random((nq, 4))), 0, vocab - 1)
    docs = rng.integers(0, len(doc_ptr) - 1, nq)
    starts, lens = doc_ptr[docs], doc_ptr[docs + 1] - doc_ptr[docs]
    pick = (rng.random((nq, 4)) * lens[:, None]).astype(np.int64)
    q[:, 4:] = doc_tok[starts[:, None] + pick]
    return q
