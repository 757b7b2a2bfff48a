"""The reference's flat-similarity scans outside the retriever on the dense kernel against the oracle,
which is pinned bit-exactly to the reference's own methods (tests/test_oracle_golden.py)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _clustered(seed, n, d, n_clusters, spread):
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((n_clusters, d)).astype(np.float32)
    x = cent[rng.integers(0, n_clusters, n)] + spread * rng.standard_normal((n, d)).astype(np.float32)
    return x.astype(np.float32)


def test_find_embedding_candidates_matches_reference_golden(golden_dir):
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.graph import find_embedding_candidates
    with open(os.path.join(golden_dir, "graph_sim_golden.json")) as f:
        golden = json.load(f)
    for c in golden["candidates"]:
        x = _clustered(c["seed"], c["n"], c["d"], 12, 0.35)
        q = (x[c["seed"]] + 0.1 * np.random.default_rng(c["seed"] + 100).standard_normal(c["d"]).astype(np.float32)).astype(np.float32)
        got = find_embedding_candidates(x, q, {f"n{i}": i for i in range(c["n_ids"])}, c["top_k"])
        if got != c["result"]:                                   # only near-ties may swap or replace entries
            sims = x @ q
            want_idx, got_idx = [int(s[1:]) for s in c["result"]], [int(s[1:]) for s in got]
            for a, b in zip(want_idx, got_idx):
                assert abs(float(sims[a]) - float(sims[b])) <= TOL * max(1.0, abs(float(sims[a]))), (c["seed"], a, b)
    assert find_embedding_candidates(None, np.zeros(4, np.float32), {}) == []


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_semantic_similarity_relations_match_oracle(golden_dir, case):
    """Every pair and rank of extract_semantic_similarity_relations; case 2 has notes with 376 neighbours above the
    threshold (more than one top-k list holds: the full-row path)."""
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.graph import semantic_similarity_relations
    from oracle import graph_sim
    with open(os.path.join(golden_dir, "graph_sim_golden.json")) as f:
        c = json.load(f)["relations"][case]
    x = _clustered(c["seed"], c["n"], c["d"], c["n_clusters"], c["spread"])
    if c["zero_row"] is not None:
        x[c["zero_row"]] = 0.0
    notes = [{"note_id": f"n{i}"} for i in range(c["n"])]
    got = semantic_similarity_relations(notes, x, c["threshold"], 0.5)
    want = graph_sim.semantic_similarity_relations(notes, x, c["threshold"], 0.5)
    assert len(want) == c["count"]
    m = graph_sim.similarity_matrix(x)
    key = lambda e: (int(e["source_id"][1:]), int(e["target_id"][1:]))
    g, w = {key(e): e for e in got}, {key(e): e for e in want}
    for pair in set(g) ^ set(w):                                  # a pair may differ only at the threshold itself
        assert abs(float(m[pair]) - c["threshold"]) <= TOL, pair
    exact_rank = 0
    for pair in set(g) & set(w):
        a, b = g[pair], w[pair]
        assert abs(a["metadata"]["cosine_similarity"] - b["metadata"]["cosine_similarity"]) <= TOL
        assert abs(a["weight"] - float(b["weight"])) <= TOL and a["relation_type"] == "semantic_similarity"
        if a["metadata"]["similarity_rank"] == b["metadata"]["similarity_rank"]:
            exact_rank += 1
        else:                                                     # ranks may differ only between near-equal similarities
            i, j = pair
            near = int((np.abs(m[i] - m[i, j]) <= 2 * TOL).sum())
            assert abs(a["metadata"]["similarity_rank"] - b["metadata"]["similarity_rank"]) < near, pair
    assert exact_rank >= 0.999 * len(set(g) & set(w))
    assert [key(e) for e in got] == sorted(g)                    # the reference's order: i ascending, then j
    assert semantic_similarity_relations(notes[:3], x, 0.7) == []    # count mismatch (:596-598)
