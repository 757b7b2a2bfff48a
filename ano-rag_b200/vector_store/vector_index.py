"""VectorIndex -- drop-in for reference vector_store/vector_index.py:9-499.

Same constructor, attributes, method names, return shapes and error convention
(never raises: logs and returns ``[]`` / ``False`` / ``""``, reference
:122-124,:202-204,:261-263).  The faiss index object is replaced by
``engine.DenseIndex`` (tcgen05 GEMM + fused top-k on the B200); there is no CPU
path -- without a CUDA device ``create_index`` fails and logs.

Differences, all stated in DESIGN.md:
  * ``b200.ivf = "exact"`` (default): every ``index_type`` is served by the exact flat kernel
    (recall 1.0); ``nlist`` / ``nprobe`` are kept as attributes, including the nlist shrink of
    ``train_index`` (:140-155), but do not approximate the search.
    ``b200.ivf = "lists"``: 'IVFFlat' is a real inverted-file index on the device -- ``train_index``
    runs k-means for ``nlist`` centroids (or ``set_centroids`` installs given ones), ``add_vectors``
    quantises rows to the centroid of largest inner product, ``search`` scans ``nprobe`` lists.
  * ``similarity_metric`` other than 'cosine' (L2) is not on the north-star path and
    is rejected by ``create_index``.
  * ``search_tensors`` is an addition: the batched tensor-returning fast path.
"""
from __future__ import annotations

import json
import os
import tempfile
from typing import Any, Dict, List, Optional

import numpy as np

from .._log import logger
from ..config import config
from .. import engine
from .._native import MAX_TOPK
from . import faiss_io

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

_SUPPORTED = ("Flat", "IVFFlat", "IVFPQ", "HNSW", "LSH")


def _cuda_available() -> bool:
    return torch is not None and torch.cuda.is_available()


class VectorIndex:
    """向量索引类 (reference docstring: builds and manages the vector index)."""

    def __init__(self, embedding_dim: int = None):
        self.embedding_dim = embedding_dim or config.get("vector_store.dimension", 768)     # :14
        self.index_type = config.get("vector_store.index_type", "IVFFlat")                 # :15
        self.similarity_metric = config.get("vector_store.similarity_metric", "cosine")    # :16
        self.use_gpu = config.get("performance.use_gpu", True) and _cuda_available()       # :17
        self.nlist = config.get("vector_store.nlist", 100)                                  # :20
        self.nprobe = config.get("vector_store.nprobe", 10)                                 # :21
        self.m = config.get("vector_store.pq_m", 8)                                         # :22
        self.storage = config.get("b200.storage", "fp32")
        self.ivf_mode = config.get("b200.ivf", "exact")                                     # 'exact' | 'lists'
        self.ivf_niter = config.get("b200.ivf_niter", 25)                                   # faiss ClusteringParameters.niter
        self.index: Optional[engine.DenseIndex] = None
        self.is_trained = False
        self.total_vectors = 0
        self._id_table = None            # row -> id given to add_vectors (None: ids are the row numbers)
        self.index_dir = config.get("storage.vector_index_path")                            # :30-39
        if not self.index_dir:
            work_dir = config.get("storage.work_dir")
            if work_dir:
                self.index_dir = os.path.join(work_dir, "vector_index")
            else:
                self.index_dir = os.path.join(tempfile.gettempdir(), "anorag_vector_index")
        os.makedirs(self.index_dir, exist_ok=True)
        self.gpu_resource = None
        logger.info(f"VectorIndex initialized: dim={self.embedding_dim}, type={self.index_type}, gpu={self.use_gpu}")

    # ------------------------------------------------------------------ create / train / add
    def create_index(self, index_type: str = None) -> bool:
        index_type = index_type or self.index_type
        try:
            logger.info(f"Creating {index_type} index with dimension {self.embedding_dim}")
            if index_type not in _SUPPORTED:
                raise ValueError(f"Unsupported index type: {index_type}")                  # :105-106
            if self.similarity_metric != "cosine":
                raise ValueError("only the cosine / inner-product metric is implemented on the B200 path")
            if self.index is not None:
                self.index.close()
            self.index = engine.DenseIndex(self.embedding_dim, storage=self.storage, normalize=True)
            self.is_trained = index_type in ("Flat", "HNSW", "LSH")                       # :80,:99,:104
            self.index_type = index_type
            logger.info(f"Index created successfully: {index_type}")
            return True
        except Exception as e:
            logger.error(f"Failed to create index: {e}")
            return False

    def train_index(self, training_vectors: np.ndarray) -> bool:
        if self.index is None:
            logger.error("Index not created yet")
            return False
        if self.is_trained:
            logger.info("Index already trained")
            return True
        try:
            if self.index_type in ("IVFFlat", "IVFPQ"):
                min_required = self.nlist * 2                                               # :139
                if len(training_vectors) < min_required:
                    new_nlist = max(1, len(training_vectors) // 2)                          # :142
                    logger.warning(f"Training data insufficient for nlist={self.nlist}, adjusting to {new_nlist}")
                    old_nlist = self.nlist
                    self.nlist = new_nlist
                    if not self.create_index(self.index_type):
                        self.nlist = old_nlist
                        return False
            if self._ivf_lists():
                self.index.train(self._as_rows(training_vectors), self.nlist, niter=self.ivf_niter)   # index.train (:158)
            self.is_trained = True      # exact search: nothing to learn
            logger.info("Index training completed")
            return True
        except Exception as e:
            logger.error(f"Failed to train index: {e}")
            return False

    def add_vectors(self, vectors, ids: Optional[np.ndarray] = None) -> bool:
        if self.index is None:
            logger.error("Index not created yet")
            return False
        if not self.is_trained and self.index_type in ("IVFFlat", "IVFPQ"):
            logger.info("Training index with provided vectors")
            if not self.train_index(vectors):
                return False
        try:
            if ids is not None and self.index_type not in ("Flat", "HNSW", "LSH"):
                # add_with_ids (:187-193): the kernels number rows in add order; ids that are not those running numbers
                # (VectorRetriever passes arange(N), retriever.py:156) are kept in a row -> id table applied to results
                ids = np.asarray(ids, dtype=np.int64).reshape(-1)
                if len(ids) != len(vectors):
                    raise ValueError("ids and vectors differ in length")
                n0 = self.index.ntotal
                running = np.arange(n0, n0 + len(ids), dtype=np.int64)
                if self._id_table is not None or not np.array_equal(ids, running):
                    base = self._id_table if self._id_table is not None else np.arange(n0, dtype=np.int64)
                    self._id_table = np.concatenate([base, ids])
            elif self._id_table is not None:
                n0 = self.index.ntotal
                self._id_table = np.concatenate([self._id_table, np.arange(n0, n0 + len(vectors), dtype=np.int64)])
            self.index.add(self._as_rows(vectors))
            self.total_vectors += len(vectors)
            logger.info(f"Added {len(vectors)} vectors to index, total: {self.total_vectors}")
            return True
        except Exception as e:
            logger.error(f"Failed to add vectors to index: {e}")
            return False

    def _ivf_lists(self) -> bool:
        return self.ivf_mode == "lists" and self.index_type == "IVFFlat"

    def set_centroids(self, centroids) -> bool:
        """Addition (BASELINE cfg5): install given coarse centroids, e.g. the topic-cluster means of
        doc/clustering.py:194, instead of training; cosine metric -> they are L2-normalised first."""
        if self.index is None:
            logger.error("Index not created yet")
            return False
        try:
            self.index.set_centroids(self._preprocess_vectors(np.asarray(centroids)))
            self.nlist = self.index.nlist
            self.is_trained = True
            return True
        except Exception as e:
            logger.error(f"Failed to set centroids: {e}")
            return False

    @staticmethod
    def _as_rows(vectors):
        """fp32 cast + contiguity of _preprocess_vectors (:268-273); the L2
        normalisation (:276-280) runs on the device inside the library."""
        if torch is not None and isinstance(vectors, torch.Tensor):
            if vectors.dtype not in (torch.float32, torch.bfloat16):
                vectors = vectors.to(torch.float32)
            return vectors.contiguous()
        vectors = np.asarray(vectors)
        if vectors.dtype != np.float32:
            vectors = vectors.astype(np.float32)
        if vectors.ndim == 1:
            vectors = vectors[None, :]
        return np.ascontiguousarray(vectors)

    # ------------------------------------------------------------------ search
    def search_tensors(self, query_vectors, top_k: int = 10):
        """(scores fp32 [nq,k], ids int64 [nq,k]); numpy in -> numpy out, CUDA in -> CUDA out."""
        if top_k > MAX_TOPK:
            # the select kernel keeps at most MAX_TOPK candidates per query (faiss has no such limit): the best MAX_TOPK are
            # returned rather than nothing (VectorRetriever.retrieve asks for 3 x top_k, retriever.py:381)
            logger.warning(f"top_k={top_k} exceeds the kernel limit; returning the best {MAX_TOPK} per query")
            top_k = MAX_TOPK
        if self._ivf_lists() and self.index.nlist > 0:
            return self.index.ivf_search(self._as_rows(query_vectors), top_k, self.nprobe)   # index.nprobe (:91)
        return self.index.search(self._as_rows(query_vectors), top_k)

    def search(self, query_vectors, top_k: int = 10, return_vectors: bool = False) -> List[Dict[str, Any]]:
        if self.index is None:
            logger.error("Index not created yet")
            return []
        if self.total_vectors == 0:
            logger.warning("Index is empty")
            return []
        try:
            scores, indices = self.search_tensors(query_vectors, top_k)                     # :220-223
            if torch is not None and isinstance(scores, torch.Tensor):
                scores, indices = scores.cpu().numpy(), indices.cpu().numpy()
            if self._id_table is not None:                                                  # ids given to add_vectors (:187-193)
                indices = np.where(indices >= 0, self._id_table[np.clip(indices, 0, len(self._id_table) - 1)], -1)
            # :227-253 -- one dict per hit, rank = position in the kernel's list, -1 slots skipped.  tolist() hands back
            # Python ints / floats, so the comprehension is the whole shaping cost (nq * k dict constructions)
            k_eff = min(top_k, indices.shape[1])
            results = [[{"index": idx, "score": sc, "rank": rank, "similarity": sc}
                        for rank, (idx, sc) in enumerate(zip(row_i, row_s)) if idx != -1]
                       for row_i, row_s in zip(indices[:, :k_eff].tolist(), scores[:, :k_eff].tolist())]
            if len(results) == 1:                                                           # :256-257
                return results[0]
            return results
        except Exception as e:
            logger.error(f"Failed to search index: {e}")
            return []

    def _preprocess_vectors(self, vectors: np.ndarray) -> np.ndarray:
        """Kept for API compatibility (:265-282).  The product path normalises on the
        device; this host version only serves external callers of the helper."""
        vectors = np.asarray(vectors)
        if vectors.dtype != np.float32:
            vectors = vectors.astype(np.float32)
        if not vectors.flags["C_CONTIGUOUS"]:
            vectors = np.ascontiguousarray(vectors)
        if self.similarity_metric == "cosine":
            norms = np.linalg.norm(vectors, axis=1, keepdims=True)
            norms = np.where(norms == 0, 1, norms)
            vectors = vectors / norms
        return vectors

    # ------------------------------------------------------------------ persistence
    def save_index(self, filename: str = None) -> str:
        """Writes faiss's own on-disk format (``faiss_io``: IndexFlatIP, or IndexIVFFlat when the index holds
        inverted lists) so the file is interchangeable with the reference's ``faiss.write_index`` (:300)."""
        if self.index is None:
            logger.error("No index to save")
            return ""
        try:
            filename = filename or f"index_{self.index_type}_{self.embedding_dim}d.faiss"   # :291
            filepath = os.path.join(self.index_dir, filename)
            rows = self.index.reconstruct(0, self.index.ntotal) if self.index.ntotal else \
                np.zeros((0, self.embedding_dim), np.float32)
            if self.index.nlist > 0:
                centroids, assign = self.index.ivf_export()
                faiss_io.write_ivf_flat(filepath, rows, centroids, assign, self.nprobe)
            else:
                faiss_io.write_flat(filepath, rows)
            metadata = {                                                                    # :303-311
                "index_type": self.index_type, "embedding_dim": self.embedding_dim,
                "similarity_metric": self.similarity_metric, "total_vectors": self.total_vectors,
                "is_trained": self.is_trained, "nlist": self.nlist, "nprobe": self.nprobe,
                "b200_storage": self.storage,                                               # extra key, ignored by the reference
            }
            with open(filepath.replace(".faiss", "_metadata.json"), "w", encoding="utf-8") as f:
                json.dump(metadata, f, ensure_ascii=False, indent=2)
            logger.info(f"Index saved to {filepath}")
            return filepath
        except Exception as e:
            logger.error(f"Failed to save index: {e}")
            return ""

    def load_index(self, filename: str) -> bool:
        """Reads an index written by ``save_index`` or by the reference's ``faiss.write_index`` (Flat / IVFFlat,
        inner-product metric); the rows go to HBM as stored (they were normalised before they were added)."""
        try:
            filepath = os.path.join(self.index_dir, filename)
            if not os.path.exists(filepath):
                logger.error(f"Index file not found: {filepath}")
                return False
            data = faiss_io.read_index(filepath)
            if data.metric != faiss_io.METRIC_INNER_PRODUCT:
                raise ValueError("only inner-product (cosine) indexes are served on the B200 path")
            metadata_file = filepath.replace(".faiss", "_metadata.json")
            self.embedding_dim = data.d
            self.total_vectors = 0
            self._id_table = None            # row -> id given to add_vectors (None: ids are the row numbers)
            if os.path.exists(metadata_file):                                               # :338-347
                with open(metadata_file, "r", encoding="utf-8") as f:
                    metadata = json.load(f)
                self.index_type = metadata.get("index_type", self.index_type)
                self.embedding_dim = metadata.get("embedding_dim", self.embedding_dim)
                self.similarity_metric = metadata.get("similarity_metric", self.similarity_metric)
                self.total_vectors = metadata.get("total_vectors", 0)
                self.is_trained = metadata.get("is_trained", True)
                self.nlist = metadata.get("nlist", self.nlist)
                self.nprobe = metadata.get("nprobe", self.nprobe)
                self.storage = metadata.get("b200_storage", self.storage)
            if self.embedding_dim != data.d:
                raise ValueError(f"metadata says dimension {self.embedding_dim}, the index file holds {data.d}")
            if self.index is not None:
                self.index.close()
            self.index = engine.DenseIndex(data.d, storage=self.storage, normalize=self.similarity_metric == "cosine")
            if data.kind == "IVFFlat" and self.ivf_mode == "lists" and data.centroids is not None:
                self.index.set_centroids(data.centroids)          # rows are re-assigned on the device while they are added
                self.nlist = self.index.nlist
            chunk = max(1, (256 << 20) // (4 * data.d))
            for lo in range(0, data.ntotal, chunk):
                self.index.add_prepared(np.ascontiguousarray(data.rows[lo: lo + chunk]))
            if not os.path.exists(metadata_file):
                self.total_vectors = data.ntotal
                self.is_trained = data.is_trained
            logger.info(f"Index loaded from {filepath}, vectors: {self.total_vectors}")
            return True
        except Exception as e:
            logger.error(f"Failed to load index: {e}")
            return False

    # ------------------------------------------------------------------ misc API
    def get_index_stats(self) -> Dict[str, Any]:
        if self.index is None:
            return {}
        stats = {"index_type": self.index_type, "embedding_dim": self.embedding_dim,
                 "similarity_metric": self.similarity_metric, "total_vectors": self.total_vectors,
                 "is_trained": self.is_trained, "use_gpu": self.use_gpu, "ntotal": self.index.ntotal}
        if self.index_type in ("IVFFlat", "IVFPQ"):
            stats["nlist"] = self.nlist
            stats["nprobe"] = self.nprobe
        return stats

    def remove_vectors(self, ids: np.ndarray) -> bool:
        if self.index is None:
            logger.error("Index not created yet")
            return False
        # IndexFlat.remove_ids compacts the storage and renumbers the rows that follow
        try:
            ids = np.unique(np.asarray(ids, dtype=np.int64))
            n = self.index.ntotal
            if self._id_table is not None:                                                  # remove_ids acts on the ids given at add time
                keep = np.flatnonzero(~np.isin(self._id_table[:n], ids))
                self._id_table = self._id_table[keep]
            else:
                keep = np.setdiff1d(np.arange(n, dtype=np.int64), ids)
            removed = n - len(keep)
            # the surviving rows never leave the device: reconstruct -> gather -> re-add as they are (already normalised)
            rows = None
            if n and len(keep):
                rows = self.index.reconstruct_device(0, n).index_select(0, torch.from_numpy(keep).to(f"cuda:{torch.cuda.current_device()}"))
            self.index.reset()
            if rows is not None and len(rows):
                self.index.add_prepared(rows.contiguous())
            self.total_vectors -= len(ids)                                                  # :404 (reference subtracts len(ids))
            logger.info(f"Removed {removed} vectors from index")
            return True
        except Exception as e:
            logger.error(f"Failed to remove vectors: {e}")
            return False

    def reset_index(self):
        if self.index is not None:
            self.index.reset()
        self.total_vectors = 0
        self._id_table = None            # row -> id given to add_vectors (None: ids are the row numbers)
        self.is_trained = False
        logger.info("Index reset completed")

    def optimize_search_params(self, query_vectors: np.ndarray, ground_truth_indices: np.ndarray,
                               target_recall: float = 0.9) -> Dict[str, Any]:
        if self.index_type not in ("IVFFlat", "IVFPQ"):
            logger.warning("Search parameter optimization only supported for IVF indices")
            return {}
        best = {"nprobe": self.nprobe, "recall": 0.0}                                       # :434
        for nprobe in [1, 5, 10, 20, 50, 100]:                                              # :437-462
            if nprobe > self.nlist:
                break
            original = self.nprobe
            self.nprobe = nprobe
            results = self.search(query_vectors, top_k=len(ground_truth_indices[0]))
            if results and isinstance(results[0], dict):
                results = [results]
            recall = self._calculate_recall(results, ground_truth_indices)
            logger.info(f"nprobe={nprobe}, recall={recall:.3f}")
            if recall > best["recall"]:
                best = {"nprobe": nprobe, "recall": recall}
            if recall >= target_recall:
                break
            self.nprobe = original
        self.nprobe = best["nprobe"]                                                        # :465-466
        logger.info(f"Optimized search params: nprobe={best['nprobe']}, recall={best['recall']:.3f}")
        return best

    def _calculate_recall(self, search_results, ground_truth: np.ndarray) -> float:
        if not search_results or len(ground_truth) == 0:
            return 0.0
        total = 0.0
        for i, query_results in enumerate(search_results):
            if i >= len(ground_truth):
                break
            retrieved = set(r["index"] for r in query_results)
            truth = set(int(x) for x in ground_truth[i])
            if truth:
                total += len(retrieved & truth) / len(truth)
        return total / min(len(search_results), len(ground_truth))

    def cleanup(self):
        if self.index is not None:
            self.index.close()
        self.index = None
        logger.info("VectorIndex cleanup completed")
