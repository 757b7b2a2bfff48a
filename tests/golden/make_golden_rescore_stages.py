#!/usr/bin/env python
"""Golden fixtures for the two remaining live fusion variants of the reference's QueryProcessor:

* ``_enhanced_hybrid_search_v2`` (query/query_processor.py:1089-1143; helpers :1235-1327)
* ``_rerank_khop_candidates``    (query/query_processor.py:1549-1635; ``_rrf_score`` :854-861,
  ``_compute_cov_cons`` :843-852)

The methods' own source is cut out of the reference file with ``ast`` and attached to a stub class (the module
itself needs LLM / graph dependencies that are absent here); the similarity providers are stubbed with preset
score lists, everything else -- section / lexical / boost helpers, RRF, coverage -- is the reference's code.
Floats are stored as float.hex().

    python tests/golden/make_golden_rescore_stages.py
"""
import ast
import json
import os
import random
import struct

REF = os.environ.get("ANORAG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

METHODS = ["_enhanced_hybrid_search_v2", "_apply_section_filtering", "_apply_lexical_fallback",
           "_satisfies_must_have_terms", "_calculate_entity_boost", "_calculate_predicate_boost",
           "_rerank_khop_candidates", "_rrf_score", "_compute_cov_cons"]


class _Log:
    """logger stub: the methods under test swallow their own exceptions and log them -- make that loud."""
    def __getattr__(self, name):
        if name == "error":
            def boom(msg, *a, **k):
                raise AssertionError(f"reference logged an error: {msg}")
            return boom
        return lambda *a, **k: None


def build_stub_class(path):
    tree = ast.parse(open(path, encoding="utf-8").read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "QueryProcessor")
    picked = []
    for node in cls.body:
        if isinstance(node, ast.FunctionDef) and node.name in METHODS:
            node.decorator_list = []
            picked.append(node)
    assert sorted(n.name for n in picked) == sorted(METHODS), [n.name for n in picked]
    stub = ast.ClassDef(name="Stub", bases=[], keywords=[], body=picked, decorator_list=[], type_params=[])
    mod = ast.fix_missing_locations(ast.Module(body=[stub], type_ignores=[]))
    ns = {"List": list, "Dict": dict, "Any": object, "Optional": object, "logger": _Log(), "config": {}}
    exec(compile(mod, path, "exec"), ns)
    return ns["Stub"]


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def hx(v):
    return float(v).hex()


def main():
    Stub = build_stub_class(os.path.join(REF, "query", "query_processor.py"))
    rng = random.Random(20261018)
    words = [f"w{i}" for i in range(30)]
    out = {"source": "reference query/query_processor.py QueryProcessor._enhanced_hybrid_search_v2 (:1089-1143) and "
                     "_rerank_khop_candidates (:1549-1635), method source executed on a stub object",
           "v2": [], "khop": []}

    def note(i):
        n = {"note_id": f"n{i}", "title": " ".join(rng.choices(words, k=rng.randint(0, 3))),
             "content": " ".join(rng.choices(words, k=rng.randint(0, 12)))}
        if rng.random() < 0.5:
            n["raw_span"] = " ".join(rng.choices(words, k=rng.randint(0, 4)))
        r = rng.random()
        if r < 0.25:
            n["section_role"] = rng.choice(["Main body", "primary topic", "appendix", "references", ""])
        return n

    # ---------------------------------------------------------------- _enhanced_hybrid_search_v2
    for ci in range(32):
        n = rng.choice([1, 2, 5, 24, 60])
        cands = [note(i) for i in range(n)]
        dense = [f32(max(0.0, rng.uniform(-0.3, 1.0))) for _ in range(n)]
        sparse = [rng.choice([0.0, rng.random() * 0.7, 1.0]) for _ in range(n)]
        if ci % 5 == 0 and n > 3:
            dense[2], sparse[2] = dense[1], sparse[1]          # equal scores keep candidate order
            cands[2] = dict(cands[1], note_id="n2")
        if ci % 6 == 1:
            dense = dense[: max(0, n - 1)]                      # short lists: missing entries are 0.0 (:1105-1106)
        if ci % 6 == 2:
            sparse = []
        must = rng.sample(words, rng.randint(1, 5)) if ci % 4 else None
        ents = rng.sample(words, rng.randint(1, 6)) if ci % 3 else None
        preds = rng.sample(words, rng.randint(1, 6)) if ci % 3 != 1 else None
        if ci == 7:
            ents = [w.upper() for w in words[:8]]               # boost cap 2.0 (:1311)
            preds = list(words[:10])                            # boost cap 1.8 (:1325)
        s = Stub()
        s.section_filtering_enabled = ci % 7 != 3
        s.section_filter_rule = "main_entity_related" if ci % 11 != 5 else "other_rule"
        s.fallback_to_lexical = ci % 2 == 0
        s.lexical_fallback_enabled = ci % 9 != 4
        s.miss_penalty = rng.choice([0.6, 0.5, 0.25])
        s.noise_threshold = rng.choice([0.20, 0.35, 0.0])
        s.listt5 = None
        s._calculate_vector_similarities = lambda q, c, v=dense: list(v)
        s._calculate_bm25_similarities = lambda q, c, v=sparse: list(v)
        work = [dict(c) for c in cands]
        res = s._enhanced_hybrid_search_v2("q", work, must, ents, preds)
        out["v2"].append({
            "settings": {"section_filtering_enabled": s.section_filtering_enabled, "section_filter_rule": s.section_filter_rule,
                         "fallback_to_lexical": s.fallback_to_lexical, "lexical_fallback_enabled": s.lexical_fallback_enabled,
                         "miss_penalty": s.miss_penalty, "noise_threshold": s.noise_threshold},
            "candidates": cands, "dense": [hx(v) for v in dense], "sparse": [hx(v) for v in sparse],
            "must_have_terms": must, "boost_entities": ents, "boost_predicates": preds,
            "order": [int(c["note_id"][1:]) for c in res],
            "final_base_score": [hx(c["final_base_score"]) for c in res],
            "all_scores": [hx(c["final_base_score"]) for c in work]})

    # ---------------------------------------------------------------- _rerank_khop_candidates
    for ci in range(32):
        n = rng.choice([1, 2, 6, 20, 75])
        cands = []
        for i in range(n):
            c = note(i)
            c["entities"] = rng.sample(words, rng.randint(0, 4)) if rng.random() < 0.8 else None
            if rng.random() < 0.7:
                c["hop_no"] = rng.choice([1, 1, 2, 2, 3, 4])
            if ci % 3 == 0 and rng.random() < 0.3:
                c["retrieval_method"] = "prf_bridge"
                if rng.random() < 0.8:
                    c["similarity_score"] = f32(rng.random())
            cands.append(c)
        if ci % 8 == 2 and n > 5:
            cands[4]["note_id"] = cands[1]["note_id"]           # duplicated note ids share one RRF entry (:1595, :858-860)
            cands[5]["note_id"] = cands[0]["note_id"]
        if ci % 8 == 6 and n > 3:
            del cands[2]["note_id"]; del cands[3]["note_id"]    # missing ids share the key None
        dense = [f32(max(0.0, rng.uniform(-0.3, 1.0))) for _ in range(n)]
        bm25 = [rng.choice([0.0, rng.random(), 1.0]) for _ in range(n)]
        focused = [f32(max(0.0, rng.uniform(-0.3, 1.0))) for _ in range(n)]
        if ci % 5 == 1 and n > 2:
            dense[1] = dense[0]; bm25[1] = bm25[0]
        path_entities = [w if rng.random() < 0.7 else w.upper() for w in rng.sample(words, rng.randint(1, 4))] if ci % 4 else None
        relation = rng.choice(["", "born in", "capital of"])
        fus = {} if ci % 5 == 0 else {"dense_weight": rng.choice([1.0, 0.8]), "bm25_weight": rng.choice([0.6, 0.4]),
                                      "focused_weight_hop2": rng.choice([0.30, 0.25, 0.05]), "rrf_lambda": rng.choice([0.2, 0.5])}
        feat = {} if ci % 5 == 0 else {"hop_decay": rng.choice([0.85, 0.7, 1.0, 0.0]), "cov_weight": rng.choice([0.10, 0.3]),
                                       "cons_weight": rng.choice([0.05, 0.2])}
        s = Stub()
        s._get_config_dict = lambda a, b, fus=fus, feat=feat: {"fusion": fus, "features": feat}[b]
        s._extract_relation_focus_phrase = lambda q, r=relation: r
        idx_of = {id(c): i for i, c in enumerate(cands)}
        work = [dict(c) for c in cands]
        pos = {id(w): i for i, w in enumerate(work)}

        def vec_sims(q, cs, dense=dense, focused=focused, pos=pos, work=work):
            if q == "q":
                return list(dense)
            return [focused[pos[id(c)]] for c in cs]            # the per-candidate "focused" query (:1612)
        s._calculate_vector_similarities = vec_sims
        s._calculate_bm25_similarities = lambda q, c, v=bm25: list(v)
        res = s._rerank_khop_candidates("q", work, path_entities)
        out["khop"].append({
            "fusion": fus, "features": feat, "candidates": cands, "dense": [hx(v) for v in dense], "bm25": [hx(v) for v in bm25],
            "focused": [hx(v) for v in focused], "path_entities": path_entities, "relation_focus": relation,
            "order": [pos[id(c)] for c in res], "final_score": [hx(c["final_score"]) for c in res]})
    with open(os.path.join(HERE, "rescore_stages_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote rescore_stages_golden.json:", len(out["v2"]), "v2 cases,", len(out["khop"]), "khop cases")


if __name__ == "__main__":
    main()
