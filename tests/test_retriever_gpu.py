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
