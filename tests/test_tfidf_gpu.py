"""The TF-IDF cosine "BM25 fallback" (vector_store/retriever.py:924-1032) on the device against the reference's own
outputs: ids, ranks and similarities bit-exact."""
import json
import os
import random

import pytest

pytestmark = pytest.mark.gpu


def _corpus(seed, n, vocab):
    rng = random.Random(seed)
    words = [f"w{i}" for i in range(vocab)] + ["the", "and", "of"]
    return [{"note_id": f"n{i}", "content": " ".join(rng.choices(words, k=rng.randint(0, 25))) + ("!" if i % 3 == 0 else "")}
            for i in range(n)]


def test_bm25_fallback_matches_reference_golden(golden_dir, tmp_path):
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.config import config
    from ano_rag_b200.vector_store import VectorRetriever
    from ano_rag_b200.vector_store.embedding_stub import HashingEmbeddingManager
    from oracle import tfidf as otf
    from sklearn.metrics.pairwise import cosine_similarity
    with open(os.path.join(golden_dir, "tfidf_golden.json")) as f:
        golden = json.load(f)
    for c in golden["cases"]:
        config.reset()
        config.update({"storage": {"vector_store_path": str(tmp_path / f"s{c['seed']}"), "vector_index_path": str(tmp_path / f"i{c['seed']}")}})
        notes = _corpus(c["seed"], c["n"], c["vocab"])
        vr = VectorRetriever(embedding_manager=HashingEmbeddingManager(32))
        assert vr.build_index(notes, save_index=False) and vr.bm25_enabled
        assert len(vr.tfidf_vectorizer.vocabulary_) == c["features"]
        ora = otf.TfidfFallback(notes)

        def same(hits, r):
            """Similarities bit-exact position by position; ids identical except inside a group of EQUAL similarities,
            whose order np.argsort (an unstable sort, :980) leaves unspecified -- the device puts the larger id first."""
            want_s = [float.fromhex(v) for v in r["sims"]]
            assert [h["retrieval_info"]["similarity"] for h in hits] == want_s, (c["seed"], r["query"])
            ids = [int(h["note_id"][1:]) for h in hits]
            if ids != r["ids"]:
                full = cosine_similarity(ora.vectorizer.transform([otf.preprocess_text(r["query"])]), ora.matrix).flatten()
                assert len(set(ids)) == len(ids) and all(float(full[i]) == s for i, s in zip(ids, want_s))
                for a in range(len(ids) - 1):                        # inside a tie group: descending id
                    assert want_s[a] != want_s[a + 1] or ids[a] > ids[a + 1]
        for r in c["results"]:                                      # one query per call, like the reference
            hits = vr._bm25_search(r["query"], c["top_k"])
            same(hits, r)
            assert [h["retrieval_info"]["rank"] for h in hits] == r["ranks"]
            assert all(h["retrieval_info"]["retrieval_method"] == "bm25_fallback" and h["retrieval_info"]["query"] == r["query"] for h in hits)
        batch = vr._bm25_search_batch([r["query"] for r in c["results"]], c["top_k"])   # all of them in one device call
        for hits, r in zip(batch, c["results"]):
            same(hits, r)
        # namespace fallback (:1004-1032): the vector hits are filtered away, the fallback's survive
        keep_fallback = lambda ns, dataset, qid: [n for n in ns if n.get("retrieval_info", {}).get("retrieval_method") == "bm25_fallback"]
        q = c["results"][0]["query"]
        res = vr.search_with_namespace_fallback([q], "d", "q1", top_k=c["top_k"], namespace_filter=keep_fallback)
        same(res[0], c["results"][0])
        vr.cleanup()
    config.reset()
