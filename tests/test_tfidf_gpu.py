"""The TF-IDF cosine "BM25 fallback" (vector_store/retriever.py:924-1032) on the device against the reference's own
outputs: ids, ranks and similarities bit-exact."""
import json
import os
import random

import numpy as np

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


@pytest.mark.parametrize("n,k", [(50_000, 1), (50_000, 100), (200_000, 1024), (3000, 1500), (700, 100)])
def test_select_kernel_equals_a_stable_sort_of_all_scores(n, k):
    """The top-k is SELECTED (radix select + a sort of the k survivors), not read off a sort of all N scores: same lists
    as ``np.argsort(similarities)[::-1]`` with a stable sort (larger doc id first among equal similarities, only
    similarities > 0, -1 beyond), including thousands of exact ties at the k-th score and k beyond the select kernel's
    list (k > 1024 takes the segmented sort)."""
    import scipy.sparse as sp
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200 import engine
    rng = np.random.default_rng(n + k)
    n_terms = 40
    # few distinct values per term -> mass ties; ~30 % of the docs hold a term
    rows, cols, vals = [], [], []
    for t in range(n_terms):
        docs = np.flatnonzero(rng.random(n) < 0.3)
        rows.append(docs); cols.append(np.full(len(docs), t)); vals.append(rng.integers(1, 6, len(docs)) / 8.0)
    m = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n_terms))
    ix = engine.TfidfIndex(m)
    # single-term queries: the similarity is one exactly rounded product, so numpy reproduces it bit for bit
    qs = sp.csr_matrix((np.array([0.5, 1.0, 0.25, 2.0]), (np.arange(4), np.array([3, 7, 11, 39]))), shape=(4, n_terms))
    s, i = ix.search(qs, k)
    dense = m.toarray()
    for r in range(4):
        t = qs[r].indices[0]
        sims = qs[r].data[0] * dense[:, t]
        order = np.argsort(sims, kind="stable")[::-1]
        order = order[sims[order] > 0][:k]
        want_i = np.full(k, -1, np.int64); want_i[:len(order)] = order
        want_s = np.zeros(k); want_s[:len(order)] = sims[order]
        assert np.array_equal(i[r], want_i) and np.array_equal(s[r], want_s), (r, n, k)
    ix.close()
