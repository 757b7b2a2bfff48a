"""VectorRetriever drop-in on the GPU with an injected deterministic encoder (the real encoder is out of scope)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_notes(n):
    topics = ["solar panels convert sunlight", "glaciers carve valleys slowly", "compilers translate source code",
              "honey bees dance to signal", "volcanoes erupt molten rock", "violins resonate through spruce"]
    return [{"note_id": f"note_{i:04d}", "content": f"{topics[i % len(topics)]} fact {i} about item{i % 17}",
             "title": f"T{i}", "entities": [f"item{i % 17}"], "paragraph_idxs": [i]} for i in range(n)]


def test_build_search_retrieve_and_hybrid(tmp_path):
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.config import config
    from ano_rag_b200.vector_store import VectorRetriever
    from ano_rag_b200.vector_store.embedding_stub import HashingEmbeddingManager
    from oracle import dense as od
    config.reset()
    config.update({"storage": {"vector_store_path": str(tmp_path / "store"), "vector_index_path": str(tmp_path / "ix")},
                   "vector_store": {"hybrid_search": {"enable_hybrid_search": True}, "similarity_threshold": 0.05},
                   "retrieval": {"candidate_pool": 30, "hybrid": {"fusion_method": "rrf"}}})
    enc = HashingEmbeddingManager(128)
    notes = make_notes(300)
    vr = VectorRetriever(embedding_manager=enc)
    assert vr.search(["anything"]) == [[]]                                   # empty index (:194-196)
    assert vr.build_index(notes) and vr.vector_index.total_vectors == 300
    queries = ["how do honey bees signal", "molten rock from volcanoes", "spruce violins"]
    res = vr.search(queries, top_k=8)
    assert len(res) == 3 and all(len(r) <= 8 for r in res)
    # same ids / similarities as the oracle's flat IP over the same embeddings
    want = od.vector_index_search(enc.encode_queries(queries), enc.encode_atomic_notes(notes), top_k=8)
    for got_q, want_q in zip(res, want):
        want_q = [w for w in want_q if w["similarity"] >= 0.05]
        assert [n["note_id"] for n in got_q] == [notes[w["index"]]["note_id"] for w in want_q]
        for n, w in zip(got_q, want_q):
            info = n["retrieval_info"]
            assert abs(info["similarity"] - w["similarity"]) <= 1e-5 and info["rank"] == w["rank"]
            assert info["retrieval_method"] == "vector_search" and n is not notes[w["index"]]      # copies (:238)
    assert "bees" in res[0][0]["content"]
    slim = vr.search(queries[:1], top_k=3, include_metadata=False)[0][0]
    assert set(slim) == {"note_id", "content", "paragraph_idxs", "retrieval_info"}
    single = vr.search_single(queries[1], top_k=4)
    assert [n["note_id"] for n in single] == [n["note_id"] for n in res[1][:4]]
    # retrieve(): boosts / penalties re-rank (:428-462)
    boosted = vr.retrieve(queries[0], top_k=5, boost_entities=["item3"], must_have_terms=["bees"])
    assert boosted and all("adjustments" in n["retrieval_info"] for n in boosted)
    sims = [n["retrieval_info"]["similarity"] for n in boosted]
    assert sims == sorted(sims, reverse=True)
    # working hybrid search: dense + BM25 + rrf on the device
    hyb = vr.hybrid_search(queries, top_k=6)
    assert len(hyb) == 3 and all(0 < len(r) <= 6 for r in hyb)
    assert all(n["retrieval_info"]["retrieval_method"] == "hybrid_search" for r in hyb for n in r)
    assert "bees" in hyb[0][0]["content"]
    # bookkeeping
    assert vr.get_note_by_id("note_0007")["content"] == notes[7]["content"]
    assert vr.add_notes(make_notes(310)[300:]) and vr.vector_index.total_vectors == 310
    assert vr.remove_notes(["note_0000", "note_0001"]) and len(vr.atomic_notes) == 308
    stats = vr.get_retrieval_stats()
    assert stats["total_notes"] == 308 and stats["index_stats"]["ntotal"] == 308
    # persistence: a fresh retriever reloads the saved index instead of re-encoding
    vr2 = VectorRetriever(embedding_manager=enc)
    assert vr2.build_index(vr.atomic_notes) and vr2.vector_index.total_vectors == 308
    assert [n["note_id"] for n in vr2.search_single(queries[2], top_k=3)] == \
           [n["note_id"] for n in vr.search_single(queries[2], top_k=3)]
    vr.cleanup(); vr2.cleanup()
    config.reset()


def test_rebuild_cli_full_incremental_and_triggers(tmp_path):
    """vector_store/rebuild_vector_index.py:127-190 on the device path: triggers (:88-119), full rebuild with faiss-format
    artefacts on disk, incremental rebuild in batches, the smoke queries of :262-272, exit codes of main()."""
    import json
    import os
    import time
    import types
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.config import config
    from ano_rag_b200.vector_store import VectorRetriever, rebuild_vector_index as rvi
    from ano_rag_b200.vector_store.embedding_stub import HashingEmbeddingManager
    config.reset()
    config.update({"storage": {"vector_store_path": str(tmp_path / "store"), "vector_index_path": str(tmp_path / "ix")},
                   "vector_store": {"similarity_threshold": 0.0},
                   "embedding_strategy": {"index_rebuilding": {"batch_size": 64, "rebuild_triggers": {"note_count_threshold": 250}}}})
    notes = make_notes(300)
    # triggers
    assert rvi.should_rebuild_index(notes) and not rvi.should_rebuild_index(notes[:100])
    old = types.SimpleNamespace(created_at=time.time() - 200 * 3600, metadata={"config_hash": rvi._calculate_config_hash()})
    fresh = types.SimpleNamespace(created_at=time.time(), metadata={"config_hash": rvi._calculate_config_hash()})
    stale_cfg = types.SimpleNamespace(created_at=time.time(), metadata={"config_hash": "other"})
    strat = lambda v: types.SimpleNamespace(current_version_id="v1", versions={"v1": v})
    assert rvi.should_rebuild_index(notes[:100], strat(old)) and rvi.should_rebuild_index(notes[:100], strat(stale_cfg))
    assert not rvi.should_rebuild_index(notes[:100], strat(fresh))
    # nothing to do -> True without touching the retriever
    assert rvi.rebuild_vector_index(notes[:100], retriever=None, embedding_strategy=strat(fresh)) is True
    # full rebuild
    vr = VectorRetriever(embedding_manager=HashingEmbeddingManager(96))
    assert rvi.rebuild_vector_index(notes, save_index=True, force_rebuild=True, retriever=vr)
    assert vr.vector_index.total_vectors == 300 and len(rvi._test_retrieval(vr)) == 3
    saved = os.listdir(str(tmp_path / "store")) + os.listdir(str(tmp_path / "ix"))
    assert any(f.endswith(".faiss") for f in saved)
    # incremental: 300 notes in batches of 64 on top of an empty retriever == the full build's results
    vr2 = VectorRetriever(embedding_manager=HashingEmbeddingManager(96))
    assert rvi.rebuild_vector_index(notes, incremental=True, retriever=vr2) and vr2.vector_index.total_vectors == 300
    a = vr.search(["honey bees dance"], top_k=5)[0]
    b = vr2.search(["honey bees dance"], top_k=5)[0]
    assert [n["note_id"] for n in a] == [n["note_id"] for n in b]
    # a failing retriever -> False, never raises
    bad = types.SimpleNamespace(build_index=lambda **k: (_ for _ in ()).throw(RuntimeError("boom")))
    assert rvi.rebuild_vector_index(notes, force_rebuild=True, retriever=bad) is False
    # main(): loaders and exit codes
    p = tmp_path / "notes.jsonl"
    p.write_text("\n".join(json.dumps(n) for n in notes[:20]))
    assert len(rvi.load_atomic_notes(str(p))) == 20
    assert rvi.main(["--data_path", str(p), "--dry_run"]) == 0
    assert rvi.main(["--data_path", str(tmp_path / "missing.json")]) == 1
    vr.cleanup(); vr2.cleanup()
