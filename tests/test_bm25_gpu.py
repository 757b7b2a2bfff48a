"""BM25 on the B200: bit-exact against the reference's golden outputs and the oracle."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def fh(x):
    return float.fromhex(x)


def test_reference_golden_simple_bm25_bit_exact(golden_dir):
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.config import config
    from ano_rag_b200.utils.bm25_search import build_bm25_corpus, bm25_scores
    config.set("b200.bm25_backend", "simple")
    with open(os.path.join(golden_dir, "bm25_simple_golden.json")) as f:
        golden = json.load(f)
    for case in golden["cases"]:
        notes = [{"content": d} for d in case["docs"]]
        corpus = build_bm25_corpus(notes, lambda n: n.get("content", ""))
        assert corpus.variant == "simple" and corpus.avgdl == fh(case["avgdl"])
        assert corpus.idf == {t: fh(v) for t, v in case["idf"].items()}
        for q in case["queries"]:
            if q["tokens"]:
                assert corpus.get_scores(q["tokens"]).tolist() == [fh(v) for v in q["raw"]], (case["name"], q["query"])
            assert bm25_scores(corpus, notes, q["query"]) == [fh(v) for v in q["normalised"]]
            half = notes[: max(1, len(notes) // 2)]
            assert bm25_scores(corpus, half, q["query"]) == [fh(v) for v in q["normalised_truncated"]]
        longer = notes + [{"content": "x"}] * 3                 # padding branch (:320-323)
        got = bm25_scores(corpus, longer, case["queries"][0]["query"])
        assert got[: len(notes)] == [fh(v) for v in case["queries"][0]["normalised"]] and got[len(notes):] == [0.0] * 3
    config.reset()


def test_okapi_matches_restated_rank_bm25():
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.utils.bm25_search import build_bm25_corpus, bm25_scores
    from oracle import bm25 as obm
    rng = random.Random(4)
    for n_docs, vocab in [(4, 30), (90, 25), (600, 300)]:
        docs = [" ".join(f"w{rng.randrange(vocab)}" for _ in range(rng.randint(1, 40))) for _ in range(n_docs)]
        if n_docs == 90:
            docs = [d + " common" for d in docs]                # df == N: negative idf -> epsilon * average_idf
        notes = [{"content": d} for d in docs]
        gpu = build_bm25_corpus(notes, lambda n: n["content"])
        ora = obm.build_bm25_corpus(notes, lambda n: n["content"], backend="okapi")
        assert gpu.variant == "okapi" and gpu.average_idf == ora.average_idf
        for _ in range(10):
            q = " ".join(f"w{rng.randrange(vocab + 3)}" for _ in range(rng.randint(1, 9)))
            if n_docs == 90:
                q += " common common"
            assert gpu.get_scores(obm.tokenize_text(q)).tolist() == ora.get_scores(obm.tokenize_text(q)).tolist()
            assert bm25_scores(gpu, notes, q) == obm.bm25_scores(ora, notes, q)
    assert bm25_scores(gpu, notes, "") == [0.0] * len(notes)     # empty query (:302-304)
    assert bm25_scores(gpu, notes, "?? !!") == [0.0] * len(notes)
    assert bm25_scores(object(), notes, "w1") == [0.0] * len(notes)


def test_okapi_reproduces_rank_bm25_readme_vector():
    """rank_bm25's README known answer ("windy London" -> [0., 0.93729472, 0.]) through the CUDA path: the only
    library-held pin for BM25Okapi available offline (utils/bm25_search.py:273,309)."""
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.utils.bm25_search import GpuBM25
    from oracle import bm25 as obm
    corpus = ["Hello there good man!", "It is quite windy in London", "How is the weather today?"]
    tok = [d.split(" ") for d in corpus]
    gpu = GpuBM25(tok, variant="okapi")
    s = gpu.get_scores("windy London".split(" "))
    assert np.round(s, 8).tolist() == [0.0, 0.93729472, 0.0]
    assert s.tolist() == obm.BM25OkapiOracle(tok).get_scores("windy London".split(" ")).tolist()
    sc, ids = gpu.search_terms(gpu.encode_queries(["windy London".split(" ")]), 2)
    assert ids[0].tolist() == [1, -1] and sc[0, 0] == s[1]
    gpu.close()


def test_degenerate_corpora_fall_back_to_simple():
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.utils.bm25_search import build_bm25_corpus, bm25_scores
    c = build_bm25_corpus([{"content": ""}, {"content": ""}], lambda n: n["content"])
    assert c.variant == "simple" and bm25_scores(c, [{}, {}], "anything") == [0.0, 0.0]
    c = build_bm25_corpus([], lambda n: n["content"])
    assert c.variant == "simple" and bm25_scores(c, [], "x") == [] and bm25_scores(c, [{}], "x") == [0.0]
    c = build_bm25_corpus([{"content": "a b"}, {"nope": 1}], lambda n: n["content"])   # text_fn raises -> []
    assert c.doc_len == [2, 0]


@pytest.mark.parametrize("variant,n,v,nq,k", [("okapi", 50_000, 4000, 300, 100), ("simple", 20_000, 500, 8, 40),
                                              ("okapi", 16_385, 50, 3, 256), ("okapi", 100, 10, 2, 100)])
@pytest.mark.parametrize("mode", ["query", "batch"])
def test_topk_and_full_scores_match_array_oracle(variant, n, v, nq, k, mode, monkeypatch):
    """Both device paths (per-query kernel, batched plane/sparse pipeline) against the restated reference."""
    monkeypatch.setenv("ARB_BM25_MODE", mode)
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.utils.bm25_search import GpuBM25
    from oracle import bm25 as obm
    rng = np.random.default_rng(n + v)
    lens = np.maximum(1, rng.poisson(25, n))
    ptr = np.zeros(n + 1, np.int64)
    ptr[1:] = np.cumsum(lens)
    p = 1.0 / np.arange(1, v + 1)
    p /= p.sum()
    tok = rng.choice(v, size=int(ptr[-1]), p=p)
    gpu = GpuBM25.from_token_ids(ptr, tok, v, variant=variant)
    ora = obm.ArrayBM25Oracle(ptr, tok, variant=variant)
    qt = rng.choice(v, size=(nq, 8), p=p).astype(np.int32)
    qt[0, 2:] = -1
    qt[1, 5] = qt[1, 0]                                          # duplicated token counts twice
    if nq > 2:
        qt[2, :] = -1                                            # nothing known -> empty result
    s, i = gpu.search_terms(qt, k)
    rows = range(nq) if nq <= 16 else list(range(8)) + rng.integers(0, nq, 8).tolist()
    full = gpu.index.scores(qt[list(rows)])
    for j, r in enumerate(rows):
        want = ora.get_scores([t for t in qt[r] if t >= 0])
        assert np.array_equal(full[j], want), f"query {r}"
        ws, wi = obm.topk_positive(want, k)
        assert np.array_equal(i[r], wi) and np.array_equal(s[r], ws), f"query {r}"
    # device-resident query terms give the same answer
    import torch
    s2, i2 = gpu.search_terms(torch.from_numpy(qt).cuda(), k)
    assert np.array_equal(s2.cpu().numpy(), s) and np.array_equal(i2.cpu().numpy(), i)
    assert gpu.index.query_postings(qt) == sum(int(ora.df[ora.term_slot[int(t)]]) for t in qt.ravel() if int(t) in ora.term_slot)
    gpu.close()


def test_reference_golden_field_weighted_bm25(golden_dir):
    """FieldWeightedBM25 (utils/bm25_search.py:66-234) against the reference's own outputs, bit-exact."""
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.utils.bm25_search import (build_field_weighted_bm25_corpus, field_weighted_bm25_scores,
                                                tokenize_text)
    with open(os.path.join(golden_dir, "bm25_field_golden.json")) as f:
        golden = json.load(f)
    for case in golden["cases"]:
        corpus = build_field_weighted_bm25_corpus(case["notes"], case["weights"])
        for q in case["queries"]:
            toks = tokenize_text(q["query"])
            if toks:
                assert corpus.get_scores(toks) == [fh(v) for v in q["raw"]], (case["name"], q["query"])
            assert field_weighted_bm25_scores(corpus, case["notes"], q["query"]) == [fh(v) for v in q["normalised"]]
        corpus.close()


@pytest.mark.parametrize("n,v,nq,k,mean_len,planes", [(200_000, 20_000, 512, 100, 40, None), (30_000, 300, 257, 10, 12, None),
                                                      (70_000, 5000, 300, 256, 30, "3"), (5_000, 40, 100, 100, 6, None),
                                                      (120_000, 2000, 200, 50, 40, "0")])
def test_batched_pipeline_equals_per_query_kernel(n, v, nq, k, mean_len, planes, monkeypatch):
    """Every query of a batch: the plane/sparse pipeline (incl. its fallback and emission paths) returns
    bit-identical scores and ids to the per-query kernel, which the tests above pin to the oracle."""
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200 import synth, _native as nat
    from ano_rag_b200.utils.bm25_search import GpuBM25
    from oracle import bm25 as obm
    if planes is not None:
        monkeypatch.setenv("ARB_BM25_MAX_PLANES", planes)
    monkeypatch.setenv("ARB_BM25_STATS", "1")
    ptr, tok = synth.bm25_corpus(n, v, mean_len, seed=n + v)
    qt = synth.bm25_queries(nq, ptr, tok, v, seed=7)
    qt[3, :] = qt[3, 0]                       # one term eight times
    qt[4, 1:] = -1
    qt[5, :] = 0                              # the densest term only (no sparse members: fallback path)
    qt[6, :4] = np.arange(4)                  # mostly plane terms
    qt[7, 2] = qt[7, 5] = qt[7, 4]            # a document's (mostly sparse) term three times, other terms in between
    qt[7, 3] = qt[7, 6]                       # and a second repeated term, interleaved
    qt[8, 4:] = qt[8, 4]                      # the same document term four times at the end
    gpu = GpuBM25.from_token_ids(ptr, tok, v, variant="okapi")
    monkeypatch.setenv("ARB_BM25_MODE", "query")
    s0, i0 = gpu.search_terms(qt, k)
    monkeypatch.setenv("ARB_BM25_MODE", "batch")
    nat.bm25_batch_stats(reset=True)
    s1, i1 = gpu.search_terms(qt, k)
    stats = nat.bm25_batch_stats()
    assert stats["queries"] == nq
    assert np.array_equal(i0, i1) and np.array_equal(s0, s1), stats
    ora = obm.ArrayBM25Oracle(ptr, tok, variant="okapi")
    for r in (0, 3, 4, 5, 6, 7, 8, nq - 1):
        ws, wi = obm.topk_positive(ora.get_scores([t for t in qt[r] if t >= 0]), k)
        assert np.array_equal(i1[r], wi) and np.array_equal(s1[r], ws), f"query {r}"
    gpu.close()


@pytest.mark.parametrize("n,v,nq,mq,k,mean_len", [(40_000, 3000, 70, 20, 1, 25), (33_333, 100_000, 128, 5, 17, 60),
                                                  (25_000, 600, 90, 32, 100, 150), (70_001, 8000, 64, 8, 250, 10)])
def test_batched_pipeline_shapes(n, v, nq, mq, k, mean_len, monkeypatch):
    """Long queries (up to the 32-token limit of the pipeline), k = 1 and k near the maximum, a vocabulary larger than
    the corpus (no planes worth having), long documents (large tf, > 4096 distinct (tf, dl) factors)."""
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200 import synth
    from ano_rag_b200.utils.bm25_search import GpuBM25
    from oracle import bm25 as obm
    ptr, tok = synth.bm25_corpus(n, v, mean_len, seed=n + mq)
    rng = np.random.default_rng(mq)
    cdf = synth.zipf_cdf(v)
    qt = np.clip(np.searchsorted(cdf, rng.random((nq, mq))), 0, v - 1).astype(np.int32)
    docs = rng.integers(0, n, nq)
    for r in range(nq):                                          # half of the tokens come from a real document
        d = tok[ptr[docs[r]]:ptr[docs[r] + 1]]
        qt[r, : mq // 2] = rng.choice(d, mq // 2)
    qt[1, 3:] = -1
    qt[2, ::2] = qt[2, 0]                                        # heavy duplication
    gpu = GpuBM25.from_token_ids(ptr, tok, v, variant="okapi")
    monkeypatch.setenv("ARB_BM25_MODE", "query")
    s0, i0 = gpu.search_terms(qt, k)
    monkeypatch.setenv("ARB_BM25_MODE", "batch")
    s1, i1 = gpu.search_terms(qt, k)
    assert np.array_equal(i0, i1) and np.array_equal(s0, s1)
    ora = obm.ArrayBM25Oracle(ptr, tok, variant="okapi")
    for r in (0, 1, 2, nq - 1):
        ws, wi = obm.topk_positive(ora.get_scores([t for t in qt[r] if t >= 0]), k)
        assert np.array_equal(i1[r], wi) and np.array_equal(s1[r], ws), f"query {r}"
    gpu.close()


def test_queries_longer_than_the_kernel_limit_keep_the_summation_order():
    """The reference scores queries of any length; the kernels take 64 tokens per call.  Longer queries are scored in
    chunks that continue the same fp64 sums on the device: bit-identical to the oracle's single pass (bm25_search.py:309)."""
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.utils.bm25_search import build_bm25_corpus, bm25_scores, build_field_weighted_bm25_corpus, field_weighted_bm25_scores
    from oracle import bm25 as obm
    rng = random.Random(9)
    docs = [" ".join(f"w{rng.randrange(120)}" for _ in range(rng.randint(5, 60))) for _ in range(700)]
    notes = [{"content": d, "title": d[:12], "entities": [d.split()[0]]} for d in docs]
    gpu = build_bm25_corpus(notes, lambda n: n["content"])
    ora = obm.build_bm25_corpus(notes, lambda n: n["content"], backend="okapi")
    for n_tok in (65, 100, 150, 200):
        q = " ".join(f"w{rng.randrange(140)}" for _ in range(n_tok))          # some tokens are outside the vocabulary
        toks = obm.tokenize_text(q)
        assert gpu.get_scores(toks).tolist() == ora.get_scores(toks).tolist(), n_tok
        got = bm25_scores(gpu, notes, q)
        assert got == obm.bm25_scores(ora, notes, q) and max(got) == 1.0
        assert bm25_scores(gpu, notes[:50], q) == obm.bm25_scores(ora, notes[:50], q)
    fw = build_field_weighted_bm25_corpus(notes)
    long_q = " ".join(f"w{rng.randrange(120)}" for _ in range(90))
    s = field_weighted_bm25_scores(fw, notes, long_q)
    assert len(s) == len(notes) and max(s) == 1.0
    assert bm25_scores(fw, notes, long_q) == s                                # FieldWeightedBM25 through the generic entry point
    fw.close(); gpu.close()
