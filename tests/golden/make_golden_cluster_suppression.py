#!/usr/bin/env python
"""Golden fixtures for the reference's near-duplicate cluster suppression
(``QueryProcessor._apply_cluster_suppression``, query/query_processor.py:1637-1697).

The method's own source is cut out of the reference file with ``ast`` and run on a stub object (the module needs LLM /
graph dependencies that are absent here).  Every candidate carries ``_embed_vec`` so the encoder (out of scope) is never
called.  Vectors are built from cluster centres so that no pairwise cosine lies within 1e-3 of the threshold: the kept
list then does not depend on fp32 rounding.  Vectors are stored as float32 hex bytes.

    python tests/golden/make_golden_cluster_suppression.py
"""
import ast
import base64
import json
import os

import numpy as np

REF = os.environ.get("ANORAG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


class _Log:
    def __getattr__(self, name):
        return lambda *a, **k: None


def build_stub(path, cfg):
    tree = ast.parse(open(path, encoding="utf-8").read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "QueryProcessor")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "_apply_cluster_suppression")
    fn.decorator_list = []
    stub = ast.ClassDef(name="Stub", bases=[], keywords=[], body=[fn], decorator_list=[], type_params=[])
    mod = ast.fix_missing_locations(ast.Module(body=[stub], type_ignores=[]))
    ns = {"List": list, "Dict": dict, "Any": object, "Optional": object, "logger": _Log()}
    exec(compile(mod, path, "exec"), ns)
    obj = ns["Stub"]()
    obj._get_config_dict = lambda *keys: cfg
    return obj


def main():
    rng = np.random.default_rng(20261018)
    cases = []
    for case_id in range(16):
        d = int(rng.choice([16, 48, 96]))
        n = int(rng.integers(1, 70))
        n_centres = int(rng.integers(1, max(2, n // 3 + 1)))
        thr = float(rng.choice([0.9, 0.8, 0.95]))
        keep_m = int(rng.choice([1, 2, 3]))
        while True:
            centres = rng.standard_normal((n_centres, d)).astype(np.float32)
            member = rng.integers(0, n_centres, n)
            spread = rng.choice([0.02, 0.15, 0.6], size=n)          # tight duplicates, loose neighbours, unrelated
            vecs = (centres[member] * rng.uniform(0.5, 2.0, (n, 1)) + spread[:, None] * rng.standard_normal((n, d))).astype(np.float32)
            if case_id % 5 == 0 and n > 2:
                vecs[1] = 0.0                                        # a zero vector: cosine 0 with everything
            u = vecs / (np.linalg.norm(vecs, axis=1, keepdims=True) + 1e-8)
            cos = u @ u.T
            if np.min(np.abs(cos[np.triu_indices(n, 1)] - thr), initial=1.0) > 1e-3:
                break
        scores = np.round(rng.uniform(0, 1, n), 3)
        if n > 3:
            scores[2] = scores[0]                                   # ties in final_score: the stable sorts decide
        cfg = {"enabled": True, "cos_threshold": thr, "keep_per_cluster": keep_m}
        cands = [{"note_id": f"n{i}", "final_score": float(scores[i]), "_embed_vec": vecs[i]} for i in range(n)]
        if case_id % 7 == 3:
            del cands[0]["final_score"]                              # .get("final_score", 0.0)
        kept = build_stub(os.path.join(REF, "query", "query_processor.py"), cfg)._apply_cluster_suppression(list(cands))
        cases.append({"dim": d, "threshold": thr, "keep_per_cluster": keep_m,
                      "final_score": [c.get("final_score") for c in cands],
                      "vecs_f32_b64": base64.b64encode(vecs.tobytes()).decode(), "kept": [c["note_id"] for c in kept]})
    disabled = build_stub(os.path.join(REF, "query", "query_processor.py"), {"enabled": "false"})._apply_cluster_suppression([{"note_id": "a"}])
    out = {"source": "reference query/query_processor.py QueryProcessor._apply_cluster_suppression (:1637-1697), method source "
                     "executed on a stub object", "cases": cases, "disabled_returns_input": [c["note_id"] for c in disabled]}
    with open(os.path.join(HERE, "cluster_suppression_golden.json"), "w") as f:
        json.dump(out, f)
    print(f"wrote {len(cases)} cases")


if __name__ == "__main__":
    main()
