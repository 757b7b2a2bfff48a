"""Host-side logic that runs without a GPU: index-build arithmetic against the oracle,
config shim, tokeniser, query encoding."""
import os
import random

import numpy as np
import pytest

import ano_rag_b200  # noqa: F401
from ano_rag_b200 import bm25_build as bb
from ano_rag_b200.config import Config
from oracle import bm25 as obm


def random_corpus(rng, n_docs, vocab, lo, hi):
    docs = [[rng.randrange(vocab) for _ in range(rng.randint(lo, hi))] for _ in range(n_docs)]
    ptr = np.zeros(n_docs + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(d) for d in docs])
    tok = np.array([t for d in docs for t in d], dtype=np.int64)
    return docs, ptr, tok


@pytest.mark.parametrize("variant", ["okapi", "simple"])
def test_build_arrays_reproduce_oracle_impacts(variant):
    rng = random.Random(11)
    for n_docs, vocab in [(5, 4), (64, 20), (700, 150)]:
        docs, ptr, tok = random_corpus(rng, n_docs, vocab, 1, 30)
        if n_docs == 64:
            tok[ptr[:-1]] = 0          # term 0 in every doc -> negative okapi idf -> epsilon floor
        arr = bb.build_arrays(ptr, tok, vocab, variant=variant)
        ora = obm.ArrayBM25Oracle(ptr, tok, variant=variant)
        assert arr.stats.avgdl == ora.avgdl
        for slot, term in enumerate(ora.vocab.tolist()):
            assert arr.stats.idf[term] == ora.idf[slot], (variant, term)
            s, e = arr.term_ptr[term], arr.term_ptr[term + 1]
            docs_o, imp_o = ora.impacts(slot)
            assert np.array_equal(arr.post_doc[s:e], docs_o)
            assert np.array_equal(arr.stats.idf[term] * arr.gtab[arr.post_gcode[s:e]], imp_o)
        if variant == "okapi":
            assert arr.stats.average_idf == ora.average_idf


def test_okapi_degenerate_corpora_raise_like_rank_bm25():
    with pytest.raises(ZeroDivisionError):
        bb.build_arrays(np.zeros(1, np.int64), np.zeros(0, np.int64), 0, variant="okapi")          # no docs
    with pytest.raises(ZeroDivisionError):
        bb.build_arrays(np.zeros(4, np.int64), np.zeros(0, np.int64), 0, variant="okapi")          # docs, no tokens
    arr = bb.build_arrays(np.zeros(4, np.int64), np.zeros(0, np.int64), 0, variant="simple")       # SimpleBM25 copes
    assert arr.stats.avgdl == 0 and len(arr.post_doc) == 0


def test_vocabulary_first_appearance_and_query_encoding():
    v = bb.Vocabulary()
    ptr, tok = v.encode_docs([["b", "a", "b"], [], ["c", "a"]])
    assert ptr.tolist() == [0, 3, 3, 5] and tok.tolist() == [0, 1, 0, 2, 1]
    assert v.encode_query(["a", "zzz", "b", "a"]) == [1, -1, 0, 1]
    assert bb.first_appearance_order(np.array([3, 1, 3, 0]), 5).tolist() == [3, 1, 0]


def test_shard_arrays_with_global_stats_match_single_index():
    rng = random.Random(2)
    docs, ptr, tok = random_corpus(rng, 300, 40, 1, 20)
    whole = bb.build_arrays(ptr, tok, 40)
    half = 150
    for lo, hi in [(0, half), (half, 300)]:
        sptr = ptr[lo:hi + 1] - ptr[lo]
        stok = tok[ptr[lo]:ptr[hi]]
        shard = bb.build_arrays(sptr, stok, 40, stats=whole.stats)
        for t in range(40):
            s, e = shard.term_ptr[t], shard.term_ptr[t + 1]
            ws, we = whole.term_ptr[t], whole.term_ptr[t + 1]
            sel = (whole.post_doc[ws:we] >= lo) & (whole.post_doc[ws:we] < hi)
            assert np.array_equal(shard.post_doc[s:e] + lo, whole.post_doc[ws:we][sel])
            assert np.array_equal(shard.gtab[shard.post_gcode[s:e]], whole.gtab[whole.post_gcode[ws:we][sel]])


def test_config_shim_dotted_get_and_effective_defaults():
    c = Config()
    assert c.get("vector_store.top_k") == 20 and c.get("retrieval.candidate_pool") == 80
    assert c.get("retrieval.hybrid.weights.bm25") == 0.5 and c.get("nope.x", 7) == 7
    c.update({"retrieval": {"hybrid": {"fusion_method": "rrf"}}})
    assert c.get("retrieval.hybrid.fusion_method") == "rrf" and c.get("retrieval.hybrid.rrf_k") == 60
    assert c.load_config()["retrieval"]["hybrid"]["weights"]["dense"] == 1.0
    c.set("b200.storage", "bf16")
    assert c.get("b200.storage") == "bf16"


def test_tokenizer_matches_oracle():
    from ano_rag_b200.utils.bm25_search import tokenize_text
    for s in ["The quick-brown fox_1 jumps!", "", "ÀÉ ünïcode 42x", "a.b,c;d"]:
        assert tokenize_text(s) == obm.tokenize_text(s)


def test_bench_layout_choice():
    """bench.py: N GPUs = R row shards x N/R query groups; 'auto' takes the fewest shards that fit the HBM budget."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    a = types.SimpleNamespace(row_shards="auto", workload="cfg3", notes=1_000_000, dim=768, hbm_budget_gb=140.0)
    assert [bench.choose_row_shards(a, w) for w in (1, 2, 4, 8)] == [1, 1, 1, 1]          # replicated corpus, split batch
    a.notes, a.dim = 100_000_000, 1024                                                    # 205 GB of rows + ~160 GB of BM25 index
    assert bench.choose_row_shards(a, 8) == 4 and bench.choose_row_shards(a, 1) == 1
    a.workload = "cfg4"
    assert bench.choose_row_shards(a, 8) == 8
    a.row_shards = "4"
    assert bench.choose_row_shards(a, 8) == 4
    a.row_shards = "3"
    with pytest.raises(SystemExit):
        bench.choose_row_shards(a, 8)
    cfg = bench.workload_config(types.SimpleNamespace(workload="cfg3", notes=10, dim=4, queries=2, topk=1, vocab=5, row_shards="2"), 8, 2)
    assert cfg["sharding"] == "2 row shard(s) x 4 query group(s)"
    ns = types.SimpleNamespace(workload="cfg3", notes=10, dim=4, queries=2, topk=1, vocab=5, row_shards="auto")
    assert bench.workload_config(ns, 8, 1) == bench.workload_config(ns, 8, 4)            # both arms print the same config at N > 1
    assert bench.workload_config(ns, 1, 1)["sharding"] == "none"


def test_bm25_scores_accepts_any_corpus_with_get_scores():
    """utils/bm25_search.py:307-314 duck-types on ``get_scores``: a foreign corpus object (rank_bm25, the reference's own
    SimpleBM25, a test dummy) is scored by itself and gets the same pad / truncate / max-normalise treatment."""
    import numpy as np
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.utils.bm25_search import bm25_scores

    class Dummy:
        def get_scores(self, tokens):
            return np.array([0.0, 2.0, 1.0, 4.0]) * len(tokens)

    docs = [{}] * 4
    assert bm25_scores(Dummy(), docs, "a b") == [0.0, 0.5, 0.25, 1.0]
    assert bm25_scores(Dummy(), docs[:2], "a b") == [0.0, 1.0]                 # truncated first, normalised after (:320-333)
    assert bm25_scores(Dummy(), docs + [{}], "a") == [0.0, 0.5, 0.25, 1.0, 0.0]
    assert bm25_scores(Dummy(), docs, "") == [0.0] * 4
    assert bm25_scores(object(), docs, "a") == [0.0] * 4

    class ListCorpus:
        def get_scores(self, tokens):
            return [0.0, 0.0]
    assert bm25_scores(ListCorpus(), [{}] * 2, "x") == [0.0, 0.0]


def test_request_coalescer_batches_concurrent_calls():
    """Concurrent single-query calls (query/query_processor.py:1478-1482, :2761-2764) share one batch call; results go back
    to their callers, a failing batch fails each of its callers, close() drains."""
    import threading
    import time
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.query.coalescer import RequestCoalescer, coalesced_search

    sizes = []

    def batch_fn(reqs):
        sizes.append(len(reqs))
        time.sleep(0.01)
        if any(r == "boom" for r in reqs):
            raise ValueError("bad batch")
        return [r * 2 for r in reqs]

    c = RequestCoalescer(batch_fn, max_batch=8, max_wait_ms=50.0)
    out = [None] * 20
    th = [threading.Thread(target=lambda i=i: out.__setitem__(i, c(i))) for i in range(20)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert out == [2 * i for i in range(20)]
    assert sum(sizes) == 20 and max(sizes) <= 8 and len(sizes) < 20 and c.requests == 20 and c.batches == len(sizes)
    f1, f2 = c.submit("boom"), c.submit(5)
    import pytest
    with pytest.raises(ValueError):
        f1.result(timeout=5)
    with pytest.raises(ValueError):
        f2.result(timeout=5)                                               # same batch, same failure
    assert c(7) == 14                                                      # the worker survives a failing batch
    c.close()
    with pytest.raises(RuntimeError):
        c.submit(1)

    class FakeRetriever:
        calls = 0

        def search(self, queries, top_k=None, similarity_threshold=None, include_metadata=True):
            FakeRetriever.calls += 1
            return [[{"note_id": q, "k": top_k}] for q in queries]

    cs = coalesced_search(FakeRetriever(), top_k=5, max_wait_ms=20.0)
    futs = [cs.submit(f"entity {i}") for i in range(6)]
    assert [f.result(timeout=5)[0]["note_id"] for f in futs] == [f"entity {i}" for i in range(6)]
    assert FakeRetriever.calls < 6
    cs.close()
