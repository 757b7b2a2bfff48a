#!/usr/bin/env python
"""Golden fixtures for the live fusion variant of the reference, QueryProcessor._hybrid_search
(query/query_processor.py:3681-3768): the method's own source is cut out of the file with `ast` and executed
with a stub `self` (the module itself needs LLM / graph dependencies that are absent here).

The reference pins numpy < 2 (requirements.txt:2), under which np.float32 similarity scalars are promoted to
fp64 when they meet Python floats; the stub therefore hands the method Python floats holding float32-precision
similarities, which is the same arithmetic.  Floats are stored as float.hex().

    python tests/golden/make_golden_rescore.py
"""
import ast
import json
import os
import random
import struct
import sys
import types

REF = os.environ.get("ANORAG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def extract_method(path, name):
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name and len(node.args.args) == 6:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            return compile(ast.fix_missing_locations(mod), path, "exec")
    raise KeyError(name)


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


class _Log:
    def __getattr__(self, _):
        return lambda *a, **k: None


def main():
    code = extract_method(os.path.join(REF, "query", "query_processor.py"), "_hybrid_search")
    ns = {"List": list, "Dict": dict, "Any": object, "logger": _Log()}
    exec(code, ns)
    hybrid_search = ns["_hybrid_search"]
    # the method imports bm25_scores from utils.bm25_search at call time: serve preset scores
    fake = types.ModuleType("utils.bm25_search")
    pkg = types.ModuleType("utils")
    sys.modules["utils"] = pkg
    sys.modules["utils.bm25_search"] = fake
    rng = random.Random(20260118)
    words = [f"w{i}" for i in range(40)]
    out = {"source": "reference query/query_processor.py:3681-3768 QueryProcessor._hybrid_search (method source executed with a stub self)",
           "cases": []}
    for ci in range(36):
        method = "linear" if ci % 2 == 0 else "rrf"
        n = rng.choice([1, 2, 7, 33, 80, 150])
        cands = [{"note_id": f"n{i}", "content": " ".join(rng.choices(words, k=rng.randint(0, 12)))} for i in range(n)]
        vec = [f32(max(0.0, rng.uniform(-0.2, 1.0))) for _ in range(n)]
        if ci % 5 == 0 and n > 2:
            vec[1] = vec[0]                                   # ties keep the candidate order (stable sort)
        bm = [rng.choice([0.0, rng.random(), 1.0]) for _ in range(n)]
        if ci % 7 == 3:
            bm = bm[: max(0, n - 2)]                          # shorter score list: missing entries count as 0.0 (:3713)
        must = rng.sample(words, rng.randint(1, 3)) if ci % 3 else None
        ents = rng.sample(words, rng.randint(1, 4)) if ci % 4 in (1, 2) else None
        preds = rng.sample(words, rng.randint(1, 3)) if ci % 4 in (2, 3) else None
        if ci == 8:
            ents = [words[0], words[0].upper()]               # the same entity twice boosts twice (:3724-3726)
        stub = types.SimpleNamespace(
            structured_logger=_Log(), hybrid_search_enabled=True, bm25_corpus=object(), fusion_method=method,
            vector_weight=0.7 if method == "linear" else 1.0, bm25_weight=0.3 if method == "linear" else 1.0,
            rrf_k=rng.choice([60, 50, 10.5]),
            _calculate_vector_similarities=lambda q, c, v=vec: list(v),
            _fallback_vector_search=lambda q, c: (_ for _ in ()).throw(RuntimeError("fallback taken")))
        fake.bm25_scores = lambda corpus, docs, query, b=bm: list(b)
        work = [dict(c) for c in cands]
        res = hybrid_search(stub, "q", work, must, ents, preds)
        out["cases"].append({
            "method": method, "vector_weight": stub.vector_weight, "bm25_weight": stub.bm25_weight, "rrf_k": stub.rrf_k,
            "contents": [c["content"] for c in cands], "vector": [float(v).hex() for v in vec],
            "bm25": [float(v).hex() for v in bm], "must_have_terms": must, "boost_entities": ents,
            "boost_predicates": preds,
            "order": [int(c["note_id"][1:]) for c in res],
            "hybrid_score": [float(c["hybrid_score"]).hex() for c in res]})
    with open(os.path.join(HERE, "rescore_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote rescore_golden.json:", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
