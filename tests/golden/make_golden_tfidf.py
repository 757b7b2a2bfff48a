#!/usr/bin/env python
"""Golden fixtures for the TF-IDF "BM25 fallback" of the reference's VectorRetriever (vector_store/retriever.py:924-1002):
_build_bm25_index, _preprocess_text and _bm25_search are cut out of the reference file with `ast` and run on a stub
object (scikit-learn is the reference's own dependency and present in this image).  Similarities as float.hex().

    python tests/golden/make_golden_tfidf.py
"""
import ast
import json
import os
import random
import re

import numpy as np
from sklearn.feature_extraction.text import TfidfVectorizer
from sklearn.metrics.pairwise import cosine_similarity

REF = os.environ.get("ANORAG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
METHODS = ["_build_bm25_index", "_preprocess_text", "_bm25_search"]


class _Log:
    def __getattr__(self, name):
        if name == "error":
            def boom(msg, *a, **k):
                raise AssertionError(f"reference logged an error: {msg}")
            return boom
        return lambda *a, **k: None


def corpus(seed, n, vocab):
    rng = random.Random(seed)
    words = [f"w{i}" for i in range(vocab)] + ["the", "and", "of"]          # a few English stop words
    return [{"note_id": f"n{i}", "content": " ".join(rng.choices(words, k=rng.randint(0, 25))) + ("!" if i % 3 == 0 else "")}
            for i in range(n)]


def main():
    path = os.path.join(REF, "vector_store", "retriever.py")
    tree = ast.parse(open(path, encoding="utf-8").read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "VectorRetriever")
    picked = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in METHODS]
    assert sorted(n.name for n in picked) == sorted(METHODS)
    stub = ast.ClassDef(name="Stub", bases=[], keywords=[], body=picked, decorator_list=[], type_params=[])
    ns = {"List": list, "Dict": dict, "Any": object, "np": np, "re": re, "logger": _Log(), "TfidfVectorizer": TfidfVectorizer,
          "cosine_similarity": cosine_similarity}
    exec(compile(ast.fix_missing_locations(ast.Module(body=[stub], type_ignores=[])), path, "exec"), ns)
    out = {"source": "reference vector_store/retriever.py:924-1002 (method source executed on a stub object); corpus(seed, n, vocab) of this script",
           "cases": []}
    for seed, n, vocab, top_k in [(1, 300, 60, 20), (2, 5000, 400, 20), (3, 40, 8, 50), (4, 9000, 30, 7)]:
        notes = corpus(seed, n, vocab)
        s = ns["Stub"]()
        s.bm25_enabled, s.atomic_notes = True, notes
        s._build_bm25_index(notes)
        rng = random.Random(seed + 50)
        queries = [" ".join(f"w{rng.randrange(vocab + 5)}" for _ in range(rng.randint(1, 6))) for _ in range(12)] + ["", "the of", "W1, w2!"]
        res = []
        for q in queries:
            hits = s._bm25_search(q, top_k)
            res.append({"query": q, "ids": [int(h["note_id"][1:]) for h in hits],
                        "sims": [float(h["retrieval_info"]["similarity"]).hex() for h in hits],
                        "ranks": [h["retrieval_info"]["rank"] for h in hits]})
        out["cases"].append({"seed": seed, "n": n, "vocab": vocab, "top_k": top_k, "features": len(s.tfidf_vectorizer.vocabulary_),
                             "results": res})
    with open(os.path.join(HERE, "tfidf_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote tfidf_golden.json:", [(c["n"], c["features"], sum(len(r["ids"]) for r in c["results"])) for c in out["cases"]])


if __name__ == "__main__":
    main()
