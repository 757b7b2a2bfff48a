"""Deterministic stand-in for the reference's EmbeddingManager (a SentenceTransformer
wrapper, vector_store/embedding_manager.py -- OUT OF SCOPE: SURVEY.md section 2 row 5).

It exists so tests, smoke() and examples can drive VectorRetriever end to end on a
box with no encoder weights.  It is NOT part of the scoring path: production code
injects the real encoder."""
import re
import zlib

import numpy as np


class HashingEmbeddingManager:
    model_name = "hashing-stub"

    def __init__(self, embedding_dim: int = 384, seed: int = 0):
        self.embedding_dim = embedding_dim
        self.seed = seed

    def _encode(self, texts):
        out = np.zeros((len(texts), self.embedding_dim), dtype=np.float32)
        for i, t in enumerate(texts):
            for w in re.findall(r"\w+", (t or "").lower()):
                h = zlib.crc32(f"{self.seed}:{w}".encode())
                rng = np.random.default_rng(h)
                out[i] += rng.standard_normal(self.embedding_dim).astype(np.float32)
        return out

    def encode_queries(self, queries):
        return self._encode(list(queries))

    def encode_atomic_notes(self, notes, include_metadata=True):
        # title || content || ENTITIES: ... like embedding_manager.py:484
        texts = []
        for n in notes:
            parts = [n.get("title", "") or "", n.get("content", "") or ""]
            ents = n.get("entities") or []
            if ents:
                parts.append("ENTITIES: " + " ".join(map(str, ents)))
            texts.append(" ".join(p for p in parts if p))
        return self._encode(texts)
