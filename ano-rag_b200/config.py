"""Minimal stand-in for the reference's global ``config`` object
(config/config_loader.py:598-606 ``get('a.b.c', default)``, :646 global instance).

Only the keys the hot-path classes read are present; values are the EFFECTIVE ones
of the shipped configuration: the reference's strict loader
drops YAML keys that are not in DEFAULT_CONFIG (config_loader.py:394-415), so
``vector_store.nlist/nprobe`` and ``performance.use_gpu`` fall back to the in-code
defaults of vector_index.py:17-22.  ``b200.*`` keys are additions of this build.
"""
from __future__ import annotations

import copy
from typing import Any, Dict

DEFAULTS: Dict[str, Any] = {
    "vector_store": {
        "dimension": 768,
        "index_type": "IVFFlat",            # vector_index.py:15 (code default)
        "similarity_metric": "cosine",      # :16
        "top_k": 20,                        # retriever.py:38
        "similarity_threshold": 0.15,       # config.yaml:95 (code default 0.5, retriever.py:39)
        "batch_size": 32,
        "bm25_fallback": {"enabled": True, "k1": 1.2, "b": 0.75},
        "retriever_enhancement": {},
    },
    "performance": {"use_gpu": True},
    "storage": {},
    "retrieval": {
        "candidate_pool": 80,
        "hybrid": {
            "enabled": True,
            "fusion_method": "linear",
            "weights": {"dense": 1.0, "bm25": 0.5, "graph": 0.35, "path": 0.1},
            "rrf_k": 60,
        },
    },
    "b200": {
        "storage": "fp32",                  # 'fp32' (exact, 3 bf16 planes) | 'bf16'
        "ivf": "exact",                     # 'exact': every index_type is the flat kernel | 'lists': IVFFlat is a real IVF
        "ivf_niter": 25,                    # k-means iterations of train_index (faiss default)
        "bm25_backend": "okapi",            # what runs wherever rank_bm25 is installed (bm25_search.py:271-275)
    },
}


class Config:
    def __init__(self, tree: Dict[str, Any] | None = None):
        self._tree = copy.deepcopy(DEFAULTS)
        if tree:
            self.update(tree)

    def get(self, key: str, default: Any = None) -> Any:
        node: Any = self._tree
        for part in key.split("."):
            if isinstance(node, dict) and part in node:
                node = node[part]
            else:
                return default
        return node

    def set(self, key: str, value: Any) -> None:
        node = self._tree
        parts = key.split(".")
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = value

    def update(self, tree: Dict[str, Any]) -> None:
        def merge(dst, src):
            for k, v in src.items():
                if isinstance(v, dict) and isinstance(dst.get(k), dict):
                    merge(dst[k], v)
                else:
                    dst[k] = copy.deepcopy(v)
        merge(self._tree, tree)

    def load_config(self) -> Dict[str, Any]:
        return copy.deepcopy(self._tree)

    def reset(self) -> None:
        self._tree = copy.deepcopy(DEFAULTS)


config = Config()
