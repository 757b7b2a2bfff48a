"""Near-duplicate cluster suppression (reference query/query_processor.py:1637-1697) on the device
against the lists the reference method kept (tests/golden/cluster_suppression_golden.json) and against the oracle."""
import base64
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cluster_suppression_golden.json")


def _cases():
    with open(GOLD) as f:
        gold = json.load(f)
    out = []
    for c in gold["cases"]:
        vecs = np.frombuffer(base64.b64decode(c["vecs_f32_b64"]), dtype=np.float32).reshape(-1, c["dim"])
        cands = []
        for i, fs in enumerate(c["final_score"]):
            e = {"note_id": f"n{i}", "_embed_vec": vecs[i].copy()}
            if fs is not None:
                e["final_score"] = fs
            cands.append(e)
        out.append((c, cands, vecs))
    return gold, out


def test_kept_lists_equal_the_reference_method():
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.query.cluster_suppression import apply_cluster_suppression, apply_cluster_suppression_batch
    gold, cases = _cases()
    for c, cands, _ in cases:
        cfg = {"enabled": True, "cos_threshold": c["threshold"], "keep_per_cluster": c["keep_per_cluster"]}
        kept = apply_cluster_suppression(list(cands), cfg)
        assert [k["note_id"] for k in kept] == c["kept"]
    # the candidate lists of several queries (same dimension) in one launch
    for dim in sorted({c["dim"] for c, _, _ in cases}):
        group = [(c, cands) for c, cands, _ in cases if c["dim"] == dim and c["threshold"] == 0.9 and c["keep_per_cluster"] == 2]
        if len(group) < 2:
            group = [(c, cands) for c, cands, _ in cases if c["dim"] == dim][:1]
        c0 = group[0][0]
        same = [(c, cands) for c, cands in group if c["threshold"] == c0["threshold"] and c["keep_per_cluster"] == c0["keep_per_cluster"]]
        outs = apply_cluster_suppression_batch([list(cands) for _, cands in same],
                                               {"cos_threshold": c0["threshold"], "keep_per_cluster": c0["keep_per_cluster"]})
        for (c, _), kept in zip(same, outs):
            assert [k["note_id"] for k in kept] == c["kept"]
    # config handling of :1643-1659
    assert [c["note_id"] for c in apply_cluster_suppression([{"note_id": "a"}], {"enabled": "false"})] == gold["disabled_returns_input"]
    assert apply_cluster_suppression([], {"enabled": True}) == []
    odd = apply_cluster_suppression(list(cases[0][1]), {"cos_threshold": "x", "keep_per_cluster": None})      # falls back to 0.90 / 2
    want = apply_cluster_suppression(list(cases[0][1]), {"cos_threshold": 0.90, "keep_per_cluster": 2})
    assert [k["note_id"] for k in odd] == [k["note_id"] for k in want]


def test_pairwise_cosines_and_embed_callable_against_the_oracle():
    import ano_rag_b200  # noqa: F401
    from ano_rag_b200.query.cluster_suppression import apply_cluster_suppression, pairwise_cosines
    from oracle import rescore as ors
    rng = np.random.default_rng(4)
    lists = [rng.standard_normal((n, 64)).astype(np.float32) for n in (1, 17, 100, 0, 33)]
    lists[1][3] = 0.0
    got = pairwise_cosines(lists)
    for v, g in zip(lists, got):
        u = v.astype(np.float64) / (np.linalg.norm(v.astype(np.float64), axis=1, keepdims=True) + 1e-8)
        assert g.shape == (len(v), len(v)) and np.allclose(g, u @ u.T, atol=2e-6)
    # vectors come from the injected encoder when a candidate has none (:1661-1667), and are cached on the candidate
    table = {f"t{i} c{i}": lists[2][i] for i in range(100)}
    table["t5 c5"] = lists[2][4] * 3.0                                      # a duplicate direction
    cands = [{"note_id": f"n{i}", "title": f"t{i}", "content": f"c{i}", "final_score": float(100 - i)} for i in range(100)]
    kept = apply_cluster_suppression(cands, {"cos_threshold": 0.9, "keep_per_cluster": 1}, embed=lambda s: table[s])
    want = ors.cluster_suppression(cands, [table[f"t{i} c{i}"] for i in range(100)], 0.9, 1)
    assert [k["note_id"] for k in kept] == [k["note_id"] for k in want] and "n5" not in [k["note_id"] for k in kept]
    assert all("_embed_vec" in c for c in cands)
