"""Oracle: the TF-IDF cosine "BM25 fallback" of the reference's VectorRetriever (TEST INFRASTRUCTURE ONLY).

Restates ``_build_bm25_index`` / ``_preprocess_text`` / ``_bm25_search`` (``vector_store/retriever.py:924-1002``) with
the same third-party calls the reference makes (scikit-learn ``TfidfVectorizer`` and ``cosine_similarity``; the
reference's requirements.txt lists scikit-learn, this image has 1.9.0).  PINNED: the methods' own source is executed
by ``tests/golden/make_golden_tfidf.py``; outputs in ``tests/golden/tfidf_golden.json``.
"""
from __future__ import annotations

import re
from typing import Any, Dict, List

import numpy as np
from sklearn.feature_extraction.text import TfidfVectorizer
from sklearn.metrics.pairwise import cosine_similarity


def preprocess_text(text: str) -> str:                               # :958-964
    text = re.sub(r"[^a-zA-Z0-9\s]", " ", text)
    text = re.sub(r"\s+", " ", text)
    return text.strip().lower()


class TfidfFallback:
    def __init__(self, atomic_notes: List[Dict[str, Any]]):
        self.atomic_notes = atomic_notes
        texts = [preprocess_text(n.get("content", "")) if n.get("content", "") else "" for n in atomic_notes]   # :928-936
        self.vectorizer = TfidfVectorizer(lowercase=True, stop_words="english", max_features=10000, ngram_range=(1, 2))
        self.matrix = self.vectorizer.fit_transform(texts)           # :948-949

    def search(self, query: str, top_k: int = 20):
        """-> [(note index, similarity)] in result order (:972-998)."""
        qv = self.vectorizer.transform([preprocess_text(query)])
        sims = cosine_similarity(qv, self.matrix).flatten()          # :977
        top = np.argsort(sims)[::-1][:top_k]                         # :980
        return [(int(i), float(sims[i])) for i in top if i < len(self.atomic_notes) and sims[i] > 0]
