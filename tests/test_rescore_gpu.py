"""Live-pipeline candidate re-scoring on the B200 against the reference's own outputs (bit-exact)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_rescore_matches_reference_golden_single_and_batched(golden_dir):
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.query import HybridRescorer
    with open(os.path.join(golden_dir, "rescore_golden.json")) as f:
        golden = json.load(f)
    by_cfg = {}
    for case in golden["cases"]:
        cands = [{"note_id": f"n{i}", "content": c} for i, c in enumerate(case["contents"])]
        vec = [float.fromhex(v) for v in case["vector"]]
        bm = [float.fromhex(v) for v in case["bm25"]]
        rs = HybridRescorer(case["method"], case["vector_weight"], case["bm25_weight"], case["rrf_k"])
        got = rs.hybrid_search([dict(c) for c in cands], vec, bm, case["must_have_terms"], case["boost_entities"], case["boost_predicates"])
        assert [int(c["note_id"][1:]) for c in got] == case["order"]
        assert [c["hybrid_score"] for c in got] == [float.fromhex(v) for v in case["hybrid_score"]]
        key = (case["method"], case["vector_weight"], case["bm25_weight"], case["rrf_k"])
        by_cfg.setdefault(key, []).append((case, cands, vec, bm))
    # ragged batches through one launch give the same answers
    for (method, vw, bw, k), group in by_cfg.items():
        rs = HybridRescorer(method, vw, bw, k)
        res = rs.rescore_batch([{"candidates": [dict(c) for c in cands], "vector_scores": vec, "bm25_scores": bm,
                                 "must_have_terms": case["must_have_terms"], "boost_entities": case["boost_entities"],
                                 "boost_predicates": case["boost_predicates"]} for case, cands, vec, bm in group])
        for (case, _, _, _), got in zip(group, res):
            assert [int(c["note_id"][1:]) for c in got] == case["order"]
            assert [c["hybrid_score"] for c in got] == [float.fromhex(v) for v in case["hybrid_score"]]


def test_rescore_large_batch_matches_oracle():
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.query import HybridRescorer
    from oracle import rescore
    rng = np.random.default_rng(3)
    words = [f"t{i}" for i in range(30)]
    for method in ("linear", "rrf"):
        batch = []
        for _ in range(300):
            n = int(rng.integers(1, 400))
            cands = [{"content": " ".join(rng.choice(words, size=int(rng.integers(0, 9))))} for _ in range(n)]
            batch.append({"candidates": cands, "vector_scores": np.float32(rng.random(n)).astype(float).tolist(),
                          "bm25_scores": np.round(rng.random(n), 2).tolist(), "must_have_terms": list(rng.choice(words, 2)),
                          "boost_entities": list(rng.choice(words, 3)), "boost_predicates": list(rng.choice(words, 2))})
        rs = HybridRescorer(method)
        want = [rescore.hybrid_search_rescore(b["vector_scores"], b["bm25_scores"], [c["content"] for c in b["candidates"]], method,
                                              rs.vector_weight, rs.bm25_weight, rs.rrf_k, b["must_have_terms"], b["boost_entities"],
                                              b["boost_predicates"]) for b in batch]
        orig = [list(b["candidates"]) for b in batch]
        got = rs.rescore_batch(batch)
        for r in range(len(batch)):
            order, score = want[r]
            assert [c["hybrid_score"] for c in got[r]] == score
            assert [id(c) for c in got[r]] == [id(orig[r][i]) for i in order]


def test_candidate_similarities_match_the_reference_loop(tmp_path):
    """VectorRetriever.calculate_vector_similarities vs the arithmetic of _calculate_vector_similarities (:3556-3581)."""
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.config import config
    from ano_rag_b200.vector_store import VectorRetriever
    from ano_rag_b200.vector_store.embedding_stub import HashingEmbeddingManager
    config.reset()
    config.update({"storage": {"vector_store_path": str(tmp_path / "store"), "vector_index_path": str(tmp_path / "ix")}})
    enc = HashingEmbeddingManager(96)
    notes = [{"note_id": f"n{i}", "content": f"topic {i % 9} about item {i} and thing {i % 5}"} for i in range(500)]
    vr = VectorRetriever(embedding_manager=enc)
    assert vr.build_index(notes)
    queries = ["item 7 thing 2", "topic 3", "unrelated words entirely"]
    pools = [notes[:40] + [{"note_id": "ghost", "content": "topic 3 about nothing"}, {"content": ""}], notes[100:103], []]
    got = vr.calculate_vector_similarities(queries, pools)
    q = np.asarray(enc.encode_queries(queries), dtype=np.float32)
    x = np.asarray(enc.encode_atomic_notes(notes), dtype=np.float32)
    assert [len(g) for g in got] == [42, 3, 0]
    for r, pool in enumerate(pools):
        for c, note in enumerate(pool):
            if note.get("note_id") in vr.note_id_to_index:
                e = x[vr.note_id_to_index[note["note_id"]]]
            elif note.get("content"):
                e = np.asarray(enc.encode_queries([note["content"]]), dtype=np.float32)[0]
            else:
                assert got[r][c] == 0.0
                continue
            want = max(0.0, float(np.dot(q[r], e) / (np.linalg.norm(q[r]) * np.linalg.norm(e))))
            assert abs(got[r][c] - want) <= 2e-6, (r, c, got[r][c], want)
    vr.cleanup()
    config.reset()


def _stage_golden(golden_dir):
    with open(os.path.join(golden_dir, "rescore_stages_golden.json")) as f:
        return json.load(f)


def test_enhanced_hybrid_search_v2_matches_reference_golden(golden_dir):
    """QueryProcessor._enhanced_hybrid_search_v2 (:1089-1143): order, kept set and every score bit-exact,
    one query per call and all cases of one setting through a single launch."""
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.query import EnhancedHybridSearchV2
    fh = float.fromhex
    cases = _stage_golden(golden_dir)["v2"]
    by_cfg = {}
    for case in cases:
        rs = EnhancedHybridSearchV2(**case["settings"])
        work = [dict(c) for c in case["candidates"]]
        got = rs.enhanced_hybrid_search_v2(work, [fh(v) for v in case["dense"]], [fh(v) for v in case["sparse"]],
                                           case["must_have_terms"], case["boost_entities"], case["boost_predicates"])
        assert [int(c["note_id"][1:]) for c in got] == case["order"]
        assert [c["final_base_score"] for c in got] == [fh(v) for v in case["final_base_score"]]
        assert [c["final_base_score"] for c in work] == [fh(v) for v in case["all_scores"]]
        by_cfg.setdefault(json.dumps(case["settings"], sort_keys=True), []).append(case)
    for key, group in by_cfg.items():
        rs = EnhancedHybridSearchV2(**json.loads(key))
        res = rs.search_batch([{"candidates": [dict(c) for c in case["candidates"]], "dense_scores": [fh(v) for v in case["dense"]],
                                "sparse_scores": [fh(v) for v in case["sparse"]], "must_have_terms": case["must_have_terms"],
                                "boost_entities": case["boost_entities"], "boost_predicates": case["boost_predicates"]}
                               for case in group])
        for case, got in zip(group, res):
            assert [int(c["note_id"][1:]) for c in got] == case["order"]
            assert [c["final_base_score"] for c in got] == [fh(v) for v in case["final_base_score"]]
    assert EnhancedHybridSearchV2().enhanced_hybrid_search_v2([], [], []) == []


def test_rerank_khop_matches_reference_golden(golden_dir):
    """QueryProcessor._rerank_khop_candidates (:1549-1635), incl. PRF ranks, duplicated / missing note ids."""
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.query import KhopReranker
    fh = float.fromhex
    cases = _stage_golden(golden_dir)["khop"]
    for case in cases:
        rr = KhopReranker(case["fusion"], case["features"])
        work = [dict(c) for c in case["candidates"]]
        has_focus = bool(KhopReranker.focus_query(case["relation_focus"], case["path_entities"]))
        got = rr.rerank_khop_candidates(work, [fh(v) for v in case["dense"]], [fh(v) for v in case["bm25"]],
                                        [fh(v) for v in case["focused"]] if has_focus else None, case["path_entities"])
        pos = {id(w): i for i, w in enumerate(work)}
        assert [pos[id(c)] for c in got] == case["order"]
        assert [c["final_score"] for c in got] == [fh(v) for v in case["final_score"]]
    assert KhopReranker().rerank_khop_candidates([], [], []) == []


def test_rescore_stages_large_batches_match_oracle():
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.query import EnhancedHybridSearchV2, KhopReranker
    from oracle import rescore
    rng = np.random.default_rng(11)
    words = [f"t{i}" for i in range(25)]

    def text(lo, hi):
        return " ".join(rng.choice(words, size=int(rng.integers(lo, hi))))
    settings = {"section_filtering_enabled": True, "section_filter_rule": "main_entity_related", "fallback_to_lexical": False,
                "lexical_fallback_enabled": True, "miss_penalty": 0.6, "noise_threshold": 0.2}
    batch, want = [], []
    for _ in range(200):
        n = int(rng.integers(1, 600))
        cands = [{"title": text(0, 3), "content": text(0, 9), "section_role": str(rng.choice(["", "main", "other"]))} for _ in range(n)]
        b = {"candidates": cands, "dense_scores": np.float32(rng.random(n)).astype(float).tolist(),
             "sparse_scores": np.round(rng.random(n), 2).tolist(), "must_have_terms": list(rng.choice(words, 3)),
             "boost_entities": list(rng.choice(words, 5)), "boost_predicates": list(rng.choice(words, 5))}
        batch.append(b)
        want.append(rescore.enhanced_hybrid_search_v2(b["dense_scores"], b["sparse_scores"], cands, settings,
                                                      b["must_have_terms"], b["boost_entities"], b["boost_predicates"]))
    orig = [list(b["candidates"]) for b in batch]
    got = EnhancedHybridSearchV2(**settings).search_batch(batch)
    for r in range(len(batch)):
        kept, score = want[r]
        assert [id(c) for c in got[r]] == [id(orig[r][i]) for i in kept]
        assert [c["final_base_score"] for c in orig[r]] == score
    batch, want = [], []
    for _ in range(150):
        n = int(rng.integers(1, 500))
        cands = [{"note_id": int(rng.integers(0, n)) if rng.random() < 0.1 else f"n{i}", "title": text(0, 3), "content": text(0, 9),
                  "entities": list(rng.choice(words, 2)), "hop_no": int(rng.integers(1, 5)),
                  **({"retrieval_method": "prf_bridge", "similarity_score": float(np.float32(rng.random()))} if rng.random() < 0.2 else {})}
                 for i in range(n)]
        b = {"candidates": cands, "dense_scores": np.float32(rng.random(n)).astype(float).tolist(),
             "bm25_scores": np.round(rng.random(n), 1).tolist(), "focused_scores": np.float32(rng.random(n)).astype(float).tolist(),
             "path_entities": list(rng.choice(words, 2))}
        batch.append(b)
        want.append(rescore.rerank_khop(b["dense_scores"], b["bm25_scores"], b["focused_scores"], cands, b["path_entities"], "x"))
    orig = [list(b["candidates"]) for b in batch]
    got = KhopReranker().rerank_batch(batch)
    for r in range(len(batch)):
        order, score = want[r]
        assert [c["final_score"] for c in orig[r]] == score
        assert [id(c) for c in got[r]] == [id(orig[r][i]) for i in order]
