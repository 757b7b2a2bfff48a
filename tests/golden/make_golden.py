#!/usr/bin/env python
"""Generate golden fixtures from the REFERENCE's own code (run in the build container).

The reference files ``utils/bm25_search.py`` and ``retrieval/hybrid_search.py`` only
need loguru/stdlib, so they are loaded BY FILE PATH from /root/reference (importing
the packages pulls jsonlines/faiss, which are absent).  Their outputs on seeded
inputs are written next to this script; floats are stored as ``float.hex()`` so the
comparison is bit-exact.  /root/reference does not exist on the GPU box, which is
why the vectors are committed.

    python tests/golden/make_golden.py
"""
import importlib.util
import json
import os
import random
import sys

REF = os.environ.get("ANORAG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def hx(x):
    return None if x is None else float(x).hex()


def make_text_corpus(rng, n_docs, vocab, min_len, max_len, zipf=True):
    words = [f"w{i}" for i in range(vocab)]
    weights = [1.0 / (i + 1) for i in range(vocab)] if zipf else None
    docs = []
    for _ in range(n_docs):
        ln = rng.randint(min_len, max_len)
        docs.append(" ".join(rng.choices(words, weights=weights, k=ln)) if ln else "")
    return docs


def main():
    from loguru import logger
    logger.remove()
    bm = load("utils/bm25_search.py", "ref_bm25_search")
    hs = load("retrieval/hybrid_search.py", "ref_hybrid_search")
    assert not bm.RANK_BM25_AVAILABLE, "fixtures below assume the SimpleBM25 fallback"

    # ---------------- BM25 (SimpleBM25 + bm25_scores wrapper) ----------------
    out = {"source": "reference utils/bm25_search.py (SimpleBM25 fallback; rank_bm25 absent)", "cases": []}
    ka_notes = [
        {"content": "The quick brown fox jumps over the lazy dog"},
        {"content": "Machine learning is a subset of artificial intelligence"},
        {"content": "Python is a popular programming language for data science"},
        {"content": "Natural language processing involves understanding human language"},
    ]  # utils/bm25_search.py:346-351
    ka_queries = ["machine learning", "python programming", "natural language", "quick fox"]  # :362-366
    corpora = [("known_answer", [n["content"] for n in ka_notes], ka_queries)]
    rng = random.Random(20261018)
    for ci, (n_docs, vocab, lo, hi) in enumerate([(37, 25, 1, 12), (120, 80, 0, 30), (256, 400, 5, 60), (64, 6, 1, 5)]):
        docs = make_text_corpus(rng, n_docs, vocab, lo, hi)
        queries = []
        for qi in range(8):
            ln = rng.randint(1, 6)
            q = rng.choices([f"w{i}" for i in range(vocab + 3)], k=ln)
            if qi == 3:
                q = q + q[:1]                       # duplicated query token
            queries.append(" ".join(q))
        queries.append("")                           # empty query -> zeros
        queries.append("zzz unknown")                # no match
        corpora.append((f"random{ci}", docs, queries))
    for name, docs, queries in corpora:
        notes = [{"content": d} for d in docs]
        corpus = bm.build_bm25_corpus(notes, lambda n: n.get("content", ""))
        case = {"name": name, "docs": docs, "queries": []}
        for q in queries:
            toks = bm.tokenize_text(q)
            raw = corpus.get_scores(toks) if toks else [0.0] * len(notes)
            norm = bm.bm25_scores(corpus, notes, q)
            short = bm.bm25_scores(corpus, notes[: max(1, len(notes) // 2)], q)   # truncation path :320-327
            case["queries"].append({"query": q, "tokens": toks,
                                    "raw": [hx(v) for v in raw],
                                    "normalised": [hx(v) for v in norm],
                                    "normalised_truncated": [hx(v) for v in short]})
        case["avgdl"] = hx(corpus.avgdl)
        case["idf"] = {t: hx(v) for t, v in sorted(corpus.idf.items())}
        out["cases"].append(case)
    with open(os.path.join(HERE, "bm25_simple_golden.json"), "w") as f:
        json.dump(out, f, indent=0)

    # ---------------- FieldWeightedBM25 (utils/bm25_search.py:66-234) ----------------
    fw = {"source": "reference utils/bm25_search.py FieldWeightedBM25 / field_weighted_bm25_scores", "cases": []}
    rng = random.Random(4242)
    for ci, (n_docs, vocab) in enumerate([(12, 8), (80, 30), (150, 120)]):
        words = [f"w{i}" for i in range(vocab)]
        notes = []
        for di in range(n_docs):
            notes.append({
                "title": " ".join(rng.choices(words, k=rng.randint(0, 4))),
                "entities": rng.choices(words, k=rng.randint(0, 3)) if di % 5 else [],
                "content": " ".join(rng.choices(words, k=rng.randint(0, 25))),
            })
        if ci == 0:
            for n in notes:
                n["entities"] = []                      # a field with avgdl == 0
        weights = None if ci != 2 else {"title": 3.0, "entities": 0.5, "content": 1.25}
        corpus = bm.build_field_weighted_bm25_corpus(notes, weights)
        case = {"name": f"fw{ci}", "notes": notes, "weights": weights, "queries": []}
        for qi in range(8):
            q = " ".join(rng.choices(words + ["zzz"], k=rng.randint(1, 6)))
            if qi == 2:
                q = q + " " + q.split()[0]
            raw = corpus.get_scores(bm.tokenize_text(q))
            norm = bm.field_weighted_bm25_scores(corpus, notes, q)
            case["queries"].append({"query": q, "raw": [hx(v) for v in raw], "normalised": [hx(v) for v in norm]})
        case["queries"].append({"query": "", "raw": [hx(0.0)] * n_docs,
                                "normalised": [hx(v) for v in bm.field_weighted_bm25_scores(corpus, notes, "")]})
        fw["cases"].append(case)
    with open(os.path.join(HERE, "bm25_field_golden.json"), "w") as f:
        json.dump(fw, f, indent=0)

    # ---------------- HybridSearcher.fuse ----------------
    fz = {"source": "reference retrieval/hybrid_search.py HybridSearcher.fuse", "cases": []}

    def run_case(cfg, dense, bm25, graph, path, name):
        searcher = hs.HybridSearcher(cfg)
        res = searcher.fuse(dense=dense, bm25=bm25, graph=graph, path=path)
        fz["cases"].append({
            "name": name, "config": cfg,
            "dense": [[i, hx(s)] for i, s in (dense or [])],
            "bm25": [[i, hx(s)] for i, s in (bm25 or [])],
            "graph": [[i, hx(s)] for i, s in (graph or [])],
            "path": [[i, hx(s)] for i, s in (path or [])],
            "result": [{"note_id": r["note_id"],
                        "scores": {k: hx(v) for k, v in r["scores"].items()},
                        "final_similarity": hx(r["final_similarity"]),
                        "tags": r["tags"]} for r in res],
        })

    w_ship = {"dense": 1.0, "bm25": 0.5, "graph": 0.35, "path": 0.1}   # shipped loader values
    for method in ("linear", "rrf"):
        cfg = {"retrieval": {"candidate_pool": 3, "hybrid": {"fusion_method": method, "rrf_k": 60,
               "weights": {"dense": 1.0, "bm25": 0.5, "graph": 0.5, "path": 0.1}}}}
        run_case(cfg, [("a", 0.9), ("b", 0.8), ("c", 0.4)], [("b", 7.0), ("c", 3.5), ("d", 1.0)], None, None,
                 f"survey_A6_{method}")
    rng = random.Random(7)
    for ci in range(48):
        method = "linear" if ci % 2 == 0 else "rrf"
        universe = rng.randint(5, 400)
        def lst(n, scale, allow_dup=False, ints=True):
            ids = [rng.randrange(universe) for _ in range(n)] if allow_dup else rng.sample(range(universe), min(n, universe))
            return [((i if ints else f"n{i}"), rng.random() * scale) for i in ids]
        ints = ci % 3 != 0
        nd, nb = rng.choice([0, 1, 5, 20, 100]), rng.choice([0, 1, 7, 40, 100])
        ng, np_ = rng.choice([0, 0, 3, 15]), rng.choice([0, 0, 2, 10])
        dense = lst(nd, 1.0, allow_dup=(ci % 7 == 0), ints=ints)
        bm25 = lst(nb, 12.0, allow_dup=(ci % 11 == 0), ints=ints)
        graph = lst(ng, 1.0, ints=ints)
        path = lst(np_, 1.0, ints=ints)
        if ci % 5 == 0 and bm25:
            bm25 = [(i, 0.0) for i, _ in bm25]            # max == 0 branch of _normalize
        if ci % 13 == 0 and dense:
            dense = [(i, -s) for i, s in dense]            # negative maximum quirk
        cfg = {"retrieval": {"candidate_pool": rng.choice([5, 50, 80, 100]),
                             "hybrid": {"fusion_method": method, "rrf_k": rng.choice([60, 50, 1]),
                                        "weights": dict(w_ship) if ci % 4 else {"dense": 1.0, "bm25": 1.0}}}}
        run_case(cfg, dense, bm25, graph, path, f"random{ci}_{method}")
    run_case({"retrieval": {"hybrid": {"enabled": False}}}, [("a", 1.0)], None, None, None, "disabled")
    run_case({}, [("a", 1.0), ("b", 0.5)], [("b", 2.0)], None, None, "bare_dict_zero_weights")
    with open(os.path.join(HERE, "fuse_golden.json"), "w") as f:
        json.dump(fz, f, indent=0)
    print("wrote", len(out["cases"]), "bm25 cases and", len(fz["cases"]), "fuse cases")


if __name__ == "__main__":
    sys.exit(main())
