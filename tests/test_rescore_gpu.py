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
