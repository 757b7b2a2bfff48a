#!/usr/bin/env python
"""Golden fixtures for the reference's flat-similarity scans outside the retriever: the methods' own source is cut out
of graph/graph_retriever.py (GraphRetriever._find_embedding_candidates, :153-170) and graph/relation_extractor.py
(RelationExtractor.extract_semantic_similarity_relations / _compute_similarity_matrix / _get_similarity_rank,
:591-633, :769-792) with `ast` and run on stub objects; inputs are regenerated from the stored seeds.

    python tests/golden/make_golden_graph_sim.py
"""
import ast
import hashlib
import json
import os
import types

import numpy as np

REF = os.environ.get("ANORAG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def stub_class(path, cls_name, methods):
    tree = ast.parse(open(path, encoding="utf-8").read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name)
    picked = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in methods]
    assert sorted(n.name for n in picked) == sorted(methods)
    for n in picked:
        n.decorator_list = []
    stub = ast.ClassDef(name="Stub", bases=[], keywords=[], body=picked, decorator_list=[], type_params=[])
    ns = {"List": list, "Dict": dict, "Any": object, "np": np, "logger": types.SimpleNamespace(warning=lambda *a, **k: None)}
    exec(compile(ast.fix_missing_locations(ast.Module(body=[stub], type_ignores=[])), path, "exec"), ns)
    return ns["Stub"]


def clustered(seed, n, d, n_clusters, spread):
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((n_clusters, d)).astype(np.float32)
    x = cent[rng.integers(0, n_clusters, n)] + spread * rng.standard_normal((n, d)).astype(np.float32)
    return x.astype(np.float32)


def main():
    GR = stub_class(os.path.join(REF, "graph", "graph_retriever.py"), "GraphRetriever", ["_find_embedding_candidates"])
    RE = stub_class(os.path.join(REF, "graph", "relation_extractor.py"), "RelationExtractor",
                    ["extract_semantic_similarity_relations", "_compute_similarity_matrix", "_get_similarity_rank"])
    out = {"source": "reference graph/graph_retriever.py:153-170 and graph/relation_extractor.py:591-633,769-792 (method source "
                     "executed on stub objects); inputs = clustered(seed, n, d, n_clusters, spread) of this script",
           "candidates": [], "relations": []}
    for seed, n, d, top_k, n_ids in [(1, 400, 48, 15, 400), (2, 10, 8, 15, 10), (3, 1000, 96, 40, 990), (4, 64, 16, 5, 64)]:
        x = clustered(seed, n, d, 12, 0.35)
        q = x[seed] + 0.1 * np.random.default_rng(seed + 100).standard_normal(d).astype(np.float32)
        g = GR()
        g.index = types.SimpleNamespace(embeddings=x, note_id_to_index={f"n{i}": i for i in range(n_ids)})
        out["candidates"].append({"seed": seed, "n": n, "d": d, "top_k": top_k, "n_ids": n_ids,
                                  "result": g._find_embedding_candidates(q.astype(np.float32), top_k)})
    for seed, n, d, n_clusters, spread, thr in [(11, 300, 32, 20, 0.30, 0.7), (12, 50, 16, 3, 0.5, 0.6), (13, 700, 64, 2, 0.45, 0.8),
                                                (14, 6, 8, 2, 0.2, 0.7)]:
        x = clustered(seed, n, d, n_clusters, spread)
        if seed == 12:
            x[7] = 0.0                                            # a zero row keeps norm 1 (:773)
        r = RE()
        r.similarity_threshold = thr
        r.relation_weights = {"semantic_similarity": 0.5}
        notes = [{"note_id": f"n{i}"} for i in range(n)]
        rel = r.extract_semantic_similarity_relations(notes, x)
        stride = max(1, len(rel) // 400)
        m = r._compute_similarity_matrix(x)
        out["relations"].append({"seed": seed, "n": n, "d": d, "n_clusters": n_clusters, "spread": spread, "threshold": thr,
                                 "zero_row": 7 if seed == 12 else None, "count": len(rel),
                                 "max_neighbours_of_one_note": int((m >= thr).sum(axis=1).max()),
                                 # every (i, j, rank) goes into the digest; every stride-th pair is stored with its similarity
                                 "digest": hashlib.sha256("".join(f"{e['source_id']},{e['target_id']},{e['metadata']['similarity_rank']};"
                                                                  for e in rel).encode()).hexdigest(),
                                 "stride": stride,
                                 "pairs": [[int(e["source_id"][1:]), int(e["target_id"][1:]), float(e["metadata"]["cosine_similarity"]).hex(),
                                            e["metadata"]["similarity_rank"], float(e["weight"]).hex()] for e in rel[::stride]]})
    with open(os.path.join(HERE, "graph_sim_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote graph_sim_golden.json:", [c["count"] for c in out["relations"]], "relations;",
          [c["max_neighbours_of_one_note"] for c in out["relations"]], "max neighbours")


if __name__ == "__main__":
    main()
