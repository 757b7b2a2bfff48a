"""The oracle against the reference's own outputs (committed golden fixtures, and the
live reference files when /root/reference is mounted).  CPU only."""
import importlib.util
import json
import os
import random

import numpy as np
import pytest

from oracle import bm25 as obm
from oracle import dense as odense
from oracle import fusion as ofu


def fh(x):
    return None if x is None else float.fromhex(x)


def load_ref(reference_dir, rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(reference_dir, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def bm25_golden(golden_dir):
    with open(os.path.join(golden_dir, "bm25_simple_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def fuse_golden(golden_dir):
    with open(os.path.join(golden_dir, "fuse_golden.json")) as f:
        return json.load(f)


def test_simple_bm25_oracle_matches_reference_golden(bm25_golden):
    for case in bm25_golden["cases"]:
        notes = [{"content": d} for d in case["docs"]]
        corpus = obm.build_bm25_corpus(notes, lambda n: n.get("content", ""), backend="simple")
        assert corpus.avgdl == fh(case["avgdl"])
        assert {t: v for t, v in corpus.idf.items()} == {t: fh(v) for t, v in case["idf"].items()}
        for q in case["queries"]:
            assert obm.tokenize_text(q["query"]) == q["tokens"]
            raw = corpus.get_scores(q["tokens"]) if q["tokens"] else [0.0] * len(notes)
            assert raw == [fh(v) for v in q["raw"]], (case["name"], q["query"])
            assert obm.bm25_scores(corpus, notes, q["query"]) == [fh(v) for v in q["normalised"]]
            half = notes[: max(1, len(notes) // 2)]
            assert obm.bm25_scores(corpus, half, q["query"]) == [fh(v) for v in q["normalised_truncated"]]


def test_known_answer_values_from_survey(bm25_golden):
    # SURVEY.md appendix A.6, recorded by running the reference file
    case = bm25_golden["cases"][0]
    assert case["name"] == "known_answer"
    raw = {q["query"]: [fh(v) for v in q["raw"]] for q in case["queries"]}
    assert raw["machine learning"] == [0.0, 2.441235179278396, 0.0, 0.0]
    assert raw["natural language"] == [0.0, 0.0, 0.6659055883108644, 2.332971991232114]
    assert fh(case["idf"]["language"]) == 0.6931471805599453


def test_okapi_restatement_known_answers():
    # PARITY UNPINNED (rank_bm25 absent): values derived from the published algorithm, SURVEY.md A.6
    docs = ["The quick brown fox jumps over the lazy dog",
            "Machine learning is a subset of artificial intelligence",
            "Python is a popular programming language for data science",
            "Natural language processing involves understanding human language"]
    c = obm.BM25OkapiOracle([obm.tokenize_text(d) for d in docs])
    assert c.average_idf == 0.7565159467742893
    assert c.get_scores(["machine", "learning"]).tolist() == [0.0, 1.7180233113841918, 0.0, 0.0]
    assert c.get_scores(["natural", "language"]).tolist() == [0.0, 0.0, 0.0, 0.9092952648057796]
    assert c.get_scores(["quick", "fox"]).tolist() == [1.627995888953579, 0.0, 0.0, 0.0]


README_CORPUS = ["Hello there good man!", "It is quite windy in London", "How is the weather today?"]


def test_okapi_oracle_reproduces_rank_bm25_readme_vector():
    """The one library-held known answer available offline: rank_bm25's README scores its three-sentence corpus
    (``doc.split(" ")`` tokens) for "windy London" as ``array([0., 0.93729472, 0.])`` -- the printed 8 decimals pin
    idf (incl. the epsilon floor: "is" has df 2 of 3 -> negative idf), avgdl and the saturation formula at
    utils/bm25_search.py:273,309."""
    c = obm.BM25OkapiOracle([d.split(" ") for d in README_CORPUS])
    s = c.get_scores("windy London".split(" "))
    assert np.round(s, 8).tolist() == [0.0, 0.93729472, 0.0]
    assert s[1].hex() == "0x1.dfe51807c1be5p-1"                  # the full fp64 value the CUDA path must reproduce
    assert c.idf["is"] == c.epsilon * c.average_idf and c.average_idf > 0


@pytest.mark.parametrize("variant", ["okapi", "simple"])
def test_array_oracle_bit_identical_to_class_oracle(variant):
    rng = random.Random(5)
    for n_docs, vocab in [(1, 3), (17, 5), (200, 60), (333, 900)]:
        docs = [[rng.randrange(vocab) for _ in range(rng.randint(1 if variant == "okapi" else 0, 25))]
                for _ in range(n_docs)]
        if variant == "okapi" and n_docs == 17:
            docs = [d + [0] for d in docs]            # df(0) == N -> negative idf -> epsilon branch
        str_docs = [[f"w{t}" for t in d] for d in docs]
        cls = obm.BM25OkapiOracle(str_docs) if variant == "okapi" else obm.SimpleBM25Oracle(str_docs)
        ptr = np.zeros(n_docs + 1, dtype=np.int64)
        ptr[1:] = np.cumsum([len(d) for d in docs])
        tok = np.array([t for d in docs for t in d], dtype=np.int64)
        arr = obm.ArrayBM25Oracle(ptr, tok, variant=variant)
        for _ in range(12):
            q = [rng.randrange(vocab + 2) for _ in range(rng.randint(1, 7))]
            a = arr.get_scores(q)
            b = cls.get_scores([f"w{t}" for t in q])
            assert np.array_equal(a, np.asarray(b, dtype=np.float64)), (variant, n_docs, q)


def _golden_args(case):
    def lst(key):
        return [(i, fh(s)) for i, s in case[key]]
    return lst("dense"), lst("bm25"), lst("graph"), lst("path")


def _golden_result(case):
    return [{"note_id": r["note_id"], "scores": {k: fh(v) for k, v in r["scores"].items()},
             "final_similarity": fh(r["final_similarity"]), "tags": r["tags"]} for r in case["result"]]


def test_fuse_oracle_matches_reference_golden(fuse_golden):
    n_exact_order = 0
    for case in fuse_golden["cases"]:
        dense, bm25, graph, path = _golden_args(case)
        got = ofu.HybridSearcherOracle(case["config"]).fuse(dense, bm25, graph, path)
        want = _golden_result(case)
        assert ofu.same_modulo_ties(got, want), case["name"]
        method = case["config"].get("retrieval", {}).get("hybrid", {}).get("fusion_method", "linear")
        if method == "rrf":
            assert got == want, case["name"]          # rrf order is deterministic in the reference
        n_exact_order += got == want
    assert n_exact_order >= len(fuse_golden["cases"]) // 2


def test_fuse_survey_values(fuse_golden):
    by_name = {c["name"]: c for c in fuse_golden["cases"]}
    lin = _golden_result(by_name["survey_A6_linear"])
    assert [(r["note_id"], r["final_similarity"]) for r in lin] == \
        [("b", 1.3888888888888888), ("a", 1.0), ("c", 0.6944444444444444)]
    rrf = _golden_result(by_name["survey_A6_rrf"])
    assert [(r["note_id"], r["final_similarity"]) for r in rrf] == \
        [("b", 0.024325753569539928), ("c", 0.02393753200204813), ("a", 0.01639344262295082)]


def test_oracle_against_live_reference_files(reference_dir):
    if reference_dir is None:
        pytest.skip("/root/reference not mounted (GPU box)")
    from loguru import logger
    logger.remove()
    bm = load_ref(reference_dir, "utils/bm25_search.py", "ref_bm25_live")
    hs = load_ref(reference_dir, "retrieval/hybrid_search.py", "ref_hs_live")
    rng = random.Random(99)
    for _ in range(6):
        n_docs, vocab = rng.randint(1, 90), rng.randint(2, 70)
        docs = [" ".join(f"w{rng.randrange(vocab)}" for _ in range(rng.randint(0, 20))) for _ in range(n_docs)]
        notes = [{"content": d} for d in docs]
        ref_c = bm.build_bm25_corpus(notes, lambda n: n["content"])
        ora_c = obm.build_bm25_corpus(notes, lambda n: n["content"], backend="simple")
        for _ in range(6):
            q = " ".join(f"w{rng.randrange(vocab + 2)}" for _ in range(rng.randint(0, 6)))
            assert bm.bm25_scores(ref_c, notes, q) == obm.bm25_scores(ora_c, notes, q)
    for ci in range(30):
        method = "linear" if ci % 2 else "rrf"
        cfg = {"retrieval": {"candidate_pool": rng.choice([3, 50, 80]),
                             "hybrid": {"fusion_method": method, "rrf_k": 60,
                                        "weights": {"dense": 1.0, "bm25": 0.5, "graph": 0.35, "path": 0.1}}}}
        uni = rng.randint(3, 150)
        mk = lambda n, sc: [(i, rng.random() * sc) for i in rng.sample(range(uni), min(n, uni))]
        args = (mk(rng.randint(0, 60), 1.0), mk(rng.randint(0, 60), 9.0), mk(rng.randint(0, 5), 1.0), mk(rng.randint(0, 4), 1.0))
        assert ofu.same_modulo_ties(ofu.HybridSearcherOracle(cfg).fuse(*args), hs.HybridSearcher(cfg).fuse(*args))


def test_dense_oracle_topk_and_shaping():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((500, 32)).astype(np.float32)
    q = rng.standard_normal((7, 32)).astype(np.float32)
    xs, qs = odense.preprocess_vectors(x), odense.preprocess_vectors(q)
    s, i = odense.flat_ip_topk(qs, xs, 10, block=128)
    full = qs @ xs.T
    for r in range(7):
        order = np.lexsort((np.arange(500), -full[r].astype(np.float64)))[:10]
        assert np.array_equal(i[r], order)
        assert np.allclose(s[r], full[r, order], atol=1e-6)
    res = odense.vector_index_search(q, x, top_k=10)
    assert len(res) == 7 and [r["rank"] for r in res[0]] == list(range(10))
    assert res[0][0]["index"] == int(i[0, 0]) and res[0][0]["similarity"] == res[0][0]["score"]
    one = odense.vector_index_search(q[:1], x, top_k=3)
    assert isinstance(one[0], dict) and len(one) == 3                  # flat list for one query (:256-257)
    s2, i2 = odense.flat_ip_topk(qs, xs[:4], 6)                        # k > N -> -1 padding
    assert (i2[:, 4:] == -1).all() and (i2[:, :4] >= 0).all()
    # zero rows survive normalisation (norm==0 -> 1)
    z = np.zeros((2, 8), np.float32)
    assert np.array_equal(odense.preprocess_vectors(z), z)


def test_bf16_helpers():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(4096).astype(np.float32)
    import torch
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(odense.bf16_round(x), want)
    hi, mid, lo = odense.split_bf16_planes(x, 3)
    assert np.array_equal((hi.astype(np.float64) + mid + lo).astype(np.float32), x)
    for p in (hi, mid, lo):
        assert np.array_equal(odense.bf16_round(p), p)


def test_rescore_oracle_matches_reference_golden(golden_dir):
    """QueryProcessor._hybrid_search arithmetic (query/query_processor.py:3681-3768), bit-exact."""
    from oracle import rescore
    with open(os.path.join(golden_dir, "rescore_golden.json")) as f:
        golden = json.load(f)
    assert len(golden["cases"]) >= 30
    for case in golden["cases"]:
        order, score = rescore.hybrid_search_rescore(
            [float.fromhex(v) for v in case["vector"]], [float.fromhex(v) for v in case["bm25"]], case["contents"],
            case["method"], case["vector_weight"], case["bm25_weight"], case["rrf_k"], case["must_have_terms"],
            case["boost_entities"], case["boost_predicates"])
        assert order == case["order"]
        assert score == [float.fromhex(v) for v in case["hybrid_score"]]


def test_rescore_stage_oracles_match_reference_golden(golden_dir):
    """_enhanced_hybrid_search_v2 (:1089-1143) and _rerank_khop_candidates (:1549-1635), bit-exact."""
    from oracle import rescore
    fh = float.fromhex
    with open(os.path.join(golden_dir, "rescore_stages_golden.json")) as f:
        golden = json.load(f)
    assert len(golden["v2"]) >= 30 and len(golden["khop"]) >= 30
    for c in golden["v2"]:
        kept, score = rescore.enhanced_hybrid_search_v2([fh(v) for v in c["dense"]], [fh(v) for v in c["sparse"]], c["candidates"],
                                                        c["settings"], c["must_have_terms"], c["boost_entities"], c["boost_predicates"])
        assert kept == c["order"] and score == [fh(v) for v in c["all_scores"]]
        assert [score[i] for i in kept] == [fh(v) for v in c["final_base_score"]]
    for c in golden["khop"]:
        order, score = rescore.rerank_khop([fh(v) for v in c["dense"]], [fh(v) for v in c["bm25"]], [fh(v) for v in c["focused"]],
                                           c["candidates"], c["path_entities"], c["relation_focus"], c["fusion"], c["features"])
        assert order == c["order"] and [score[i] for i in order] == [fh(v) for v in c["final_score"]]


def test_ivf_oracle_is_consistent_with_flat_oracle():
    """IVF restatement: probing every list equals the flat search; fewer probes only ever drop rows; padding is -1."""
    from oracle import dense as od
    rng = np.random.default_rng(11)
    x = od.preprocess_vectors(rng.standard_normal((600, 24)).astype(np.float32))
    q = od.preprocess_vectors(rng.standard_normal((9, 24)).astype(np.float32))
    cent = od.preprocess_vectors(rng.standard_normal((13, 24)).astype(np.float32))
    assign = od.ivf_assign(x, cent)
    assert assign.shape == (600,) and assign.min() >= 0 and assign.max() < 13
    assert np.array_equal(assign, np.argmax(x @ cent.T, axis=1))
    fs, fi = od.flat_ip_topk(q, x, 10)
    s, i, probes = od.ivf_search(q, x, cent, assign, 10, 13)
    assert np.array_equal(i, fi) and np.allclose(s, fs, atol=1e-6) and probes.shape == (9, 13)   # BLAS gemv vs gemm order
    s3, i3, p3 = od.ivf_search(q, x, cent, assign, 10, 3)
    for r in range(9):
        allowed = set(np.flatnonzero(np.isin(assign, p3[r])).tolist())
        got = [int(v) for v in i3[r] if v >= 0]
        assert set(got) <= allowed and got == [int(v) for v in fi_restricted(q[r], x, sorted(allowed), 10)]
    tiny = od.ivf_search(q[:1], x[:3], cent, assign[:3], 10, 13)
    assert (tiny[1][0, 3:] == -1).all() and (tiny[0][0, 3:] == od.LOWEST_F32).all()


def fi_restricted(qv, x, rows, k):
    rows = np.asarray(rows)
    s = x[rows] @ qv
    order = np.lexsort((rows, -s.astype(np.float64)))[:k]
    return rows[order]


def _clustered(seed, n, d, n_clusters, spread):
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((n_clusters, d)).astype(np.float32)
    x = cent[rng.integers(0, n_clusters, n)] + spread * rng.standard_normal((n, d)).astype(np.float32)
    return x.astype(np.float32)


def test_graph_similarity_oracle_matches_reference_golden(golden_dir):
    """_find_embedding_candidates (graph/graph_retriever.py:153-170) and extract_semantic_similarity_relations
    (graph/relation_extractor.py:591-633) restated; same numpy arithmetic, so bit-exact."""
    import hashlib
    from oracle import graph_sim
    with open(os.path.join(golden_dir, "graph_sim_golden.json")) as f:
        golden = json.load(f)
    for c in golden["candidates"]:
        x = _clustered(c["seed"], c["n"], c["d"], 12, 0.35)
        q = (x[c["seed"]] + 0.1 * np.random.default_rng(c["seed"] + 100).standard_normal(c["d"]).astype(np.float32)).astype(np.float32)
        got = graph_sim.find_embedding_candidates(x, q, {f"n{i}": i for i in range(c["n_ids"])}, c["top_k"])
        assert got == c["result"]
    for c in golden["relations"]:
        x = _clustered(c["seed"], c["n"], c["d"], c["n_clusters"], c["spread"])
        if c["zero_row"] is not None:
            x[c["zero_row"]] = 0.0
        rel = graph_sim.semantic_similarity_relations([{"note_id": f"n{i}"} for i in range(c["n"])], x, c["threshold"], 0.5)
        assert len(rel) == c["count"]
        text = "".join(f"{e['source_id']},{e['target_id']},{e['metadata']['similarity_rank']};" for e in rel)
        assert hashlib.sha256(text.encode()).hexdigest() == c["digest"]
        for e, want in zip(rel[::c["stride"]], c["pairs"]):
            assert [int(e["source_id"][1:]), int(e["target_id"][1:])] == want[:2] and e["metadata"]["similarity_rank"] == want[3]
            assert e["metadata"]["cosine_similarity"] == float.fromhex(want[2]) and float(e["weight"]) == float.fromhex(want[4])


def _tfidf_corpus(seed, n, vocab):
    import random
    rng = random.Random(seed)
    words = [f"w{i}" for i in range(vocab)] + ["the", "and", "of"]
    return [{"note_id": f"n{i}", "content": " ".join(rng.choices(words, k=rng.randint(0, 25))) + ("!" if i % 3 == 0 else "")}
            for i in range(n)]


def test_tfidf_fallback_oracle_matches_reference_golden(golden_dir):
    """VectorRetriever._bm25_search (vector_store/retriever.py:966-1002): ids, ranks and similarities bit-exact."""
    from oracle import tfidf
    with open(os.path.join(golden_dir, "tfidf_golden.json")) as f:
        golden = json.load(f)
    for c in golden["cases"][:3]:
        fb = tfidf.TfidfFallback(_tfidf_corpus(c["seed"], c["n"], c["vocab"]))
        assert len(fb.vectorizer.vocabulary_) == c["features"]
        for r in c["results"]:
            got = fb.search(r["query"], c["top_k"])
            assert [i for i, _ in got] == r["ids"] and [s for _, s in got] == [float.fromhex(v) for v in r["sims"]]


def _cluster_cases(golden_dir):
    import base64
    with open(os.path.join(golden_dir, "cluster_suppression_golden.json")) as f:
        gold = json.load(f)
    for c in gold["cases"]:
        vecs = np.frombuffer(base64.b64decode(c["vecs_f32_b64"]), dtype=np.float32).reshape(-1, c["dim"])
        cands = []
        for i, fs in enumerate(c["final_score"]):
            e = {"note_id": f"n{i}", "_embed_vec": vecs[i].copy()}
            if fs is not None:
                e["final_score"] = fs
            cands.append(e)
        yield c, cands, vecs
    return gold


def test_cluster_suppression_oracle_matches_reference_golden(golden_dir):
    """query/query_processor.py:1637-1697 restated in oracle/rescore.py against lists kept by the reference method itself."""
    from oracle import rescore as ors
    n = 0
    for c, cands, vecs in _cluster_cases(golden_dir):
        kept = ors.cluster_suppression(cands, list(vecs), c["threshold"], c["keep_per_cluster"])
        assert [k["note_id"] for k in kept] == c["kept"]
        n += 1
    assert n == 16
