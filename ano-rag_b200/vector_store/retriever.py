"""VectorRetriever -- drop-in for reference vector_store/retriever.py:29-1034 on the
candidate-scoring path: build_index (:118-184), search (:186-272), search_single
(:331-337), retrieve (:339-512), hybrid_search (:274-329), note bookkeeping
(:514-672), persistence (:680-749), stats (:751-765).

The sentence encoder (EmbeddingManager, out of scope: SURVEY.md section 2 row 5) is
injected; anything with ``embedding_dim``, ``encode_queries(list[str])`` and
``encode_atomic_notes(notes, include_metadata=True)`` works, including the
reference's own class.  Embeddings may come back as numpy arrays (reference) or as
CUDA tensors (stay on the device).

Additions: ``search_tensors`` (batched tensor fast path) and a working
``hybrid_search`` -- the reference delegates to HybridSearcher.search/build_index,
which do not exist (retriever.py:170,288 vs hybrid_search.py:13-103), so it always
degrades to ``search``; here, when ``vector_store.hybrid_search.enable_hybrid_search``
is set, dense top-k + BM25 top-k + HybridSearcher fusion run on the device.
"""
from __future__ import annotations

import json
import os
import tempfile
from typing import Any, Callable, Dict, List, Optional

import numpy as np

from .._log import logger
from ..config import config
from .vector_index import VectorIndex
from .. import engine

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _note_text(note: Dict[str, Any]) -> str:
    content = note.get("content", "")
    if isinstance(content, dict):
        content = content.get("text", "") or content.get("content", "") or str(content)
    elif not isinstance(content, str):
        content = str(content)
    return content


class VectorRetriever:
    def __init__(self, embedding_manager=None):
        if embedding_manager is None:
            # the reference constructs its singleton encoder here (retriever.py:34)
            from vector_store.embedding_manager import EmbeddingManager  # type: ignore  # reference package
            embedding_manager = EmbeddingManager()
        self.embedding_manager = embedding_manager
        self.vector_index = VectorIndex(self.embedding_manager.embedding_dim)                # :35
        self.top_k = config.get("vector_store.top_k", 20)                                    # :38
        self.similarity_threshold = config.get("vector_store.similarity_threshold", 0.5)     # :39
        self.batch_size = config.get("vector_store.batch_size", 32)
        enh = config.get("vector_store.retriever_enhancement", {}) or {}
        self.default_topk_multiplier = enh.get("topk_multiplier", 3.0)                       # :44-48
        self.must_have_terms_penalty = enh.get("must_have_terms_penalty", 0.6)
        self.entity_boost_factor = enh.get("entity_boost_factor", 1.2)
        self.predicate_boost_factor = enh.get("predicate_boost_factor", 1.15)
        self.enable_filter_logging = enh.get("enable_filter_logging", True)
        self.atomic_notes: List[Dict[str, Any]] = []
        self.note_embeddings = None
        self.note_id_to_index: Dict[Any, int] = {}
        self.index_to_note_id: Dict[int, Any] = {}
        default_path = config.get("storage.vector_store_path")
        if not default_path:
            default_path = os.path.join(tempfile.gettempdir(), "anorag_vector_store")
        self.data_dir = default_path
        os.makedirs(self.data_dir, exist_ok=True)
        self.hybrid_searcher = None
        self.hybrid_config = config.get("vector_store.hybrid_search", {}) or {}              # :85-87
        self.enable_hybrid_search = bool(self.hybrid_config.get("enable_hybrid_search", False))
        self.bm25_corpus = None
        self.bm25_enabled = config.get("vector_store.bm25_fallback.enabled", True)           # :77
        self.tfidf_vectorizer = None
        self.tfidf_matrix = None
        self.processed_texts: List[str] = []
        self._tfidf_index = None
        if self.enable_hybrid_search:
            try:
                from ..retrieval.hybrid_search import create_hybrid_searcher
                self.hybrid_searcher = create_hybrid_searcher(config)
                logger.info("Hybrid searcher initialized")
            except Exception as e:
                logger.warning(f"Failed to initialize hybrid searcher: {e}")
                self.enable_hybrid_search = False
        logger.info("VectorRetriever initialized")

    # ------------------------------------------------------------------ build
    def build_index(self, atomic_notes: List[Dict[str, Any]], force_rebuild: bool = False,
                    save_index: bool = True) -> bool:
        if not atomic_notes:
            logger.warning("No atomic notes provided for indexing")
            return False
        try:
            logger.info(f"Building vector index for {len(atomic_notes)} atomic notes")
            if not force_rebuild and self._can_load_existing_index(atomic_notes):
                logger.info("Loading existing index")
                return True
            self.atomic_notes = atomic_notes
            self._build_id_mappings()
            self.note_embeddings = self.embedding_manager.encode_atomic_notes(atomic_notes, include_metadata=True)
            if self._size(self.note_embeddings) == 0:
                logger.error("Failed to generate embeddings")
                return False
            if not self.vector_index.create_index():
                logger.error("Failed to create vector index")
                return False
            self.vector_index.total_vectors = 0
            note_ids = np.arange(len(atomic_notes), dtype=np.int64)                          # :156
            if not self.vector_index.add_vectors(self.note_embeddings, note_ids):
                logger.error("Failed to add vectors to index")
                return False
            if self.bm25_enabled:                                                            # :162-164
                logger.info("Building BM25 index for fallback retrieval")
                self._build_bm25_index(atomic_notes)
            if self.enable_hybrid_search and self.hybrid_searcher:
                try:
                    from ..utils.bm25_search import build_bm25_corpus
                    self.bm25_corpus = build_bm25_corpus(atomic_notes, _note_text)
                except Exception as e:
                    logger.warning(f"Failed to build hybrid search index: {e}")
                    self.enable_hybrid_search = False
            if save_index:
                self._save_index_data()
            logger.info(f"Vector index built successfully with {len(atomic_notes)} notes")
            return True
        except Exception as e:
            logger.error(f"Failed to build vector index: {e}")
            return False

    @staticmethod
    def _size(x) -> int:
        if x is None:
            return 0
        if torch is not None and isinstance(x, torch.Tensor):
            return x.numel()
        return int(np.asarray(x).size)

    # ------------------------------------------------------------------ search
    def search_tensors(self, queries: List[str], top_k: Optional[int] = None):
        """(scores [nq,k] fp32, ids [nq,k] int64) for a batch of query strings."""
        emb = self.embedding_manager.encode_queries(queries)
        return self.vector_index.search_tensors(emb, top_k or self.top_k)

    def calculate_vector_similarities(self, queries: List[str], notes_pools: List[List[Dict[str, Any]]]) -> List[List[float]]:
        """Batched form of QueryProcessor._calculate_vector_similarities (reference query/query_processor.py:3543-3589):
        cosine similarity of every query with every note of its candidate pool, clamped at 0.  Notes known to the
        index are scored on the device from their stored embeddings (one gather kernel for the whole batch); a note
        without a stored embedding is encoded from its content like the reference does (:3572-3581); 0.0 otherwise.
        Never raises (:3587-3589)."""
        try:
            q_emb = self.embedding_manager.encode_queries(list(queries))
            nq = len(queries)
            nc = max((len(p) for p in notes_pools), default=0)
            out = [[0.0] * len(p) for p in notes_pools]
            if nq == 0 or nc == 0:
                return out
            cand = np.full((nq, nc), -1, dtype=np.int64)
            dynamic = []
            total = self.vector_index.total_vectors
            for r, pool in enumerate(notes_pools):
                for c, note in enumerate(pool):
                    nid = note.get("note_id")
                    idx = self.note_id_to_index.get(nid) if nid else None
                    if idx is not None and idx < total:
                        cand[r, c] = idx
                    elif idx is None and note.get("content", ""):
                        dynamic.append((r, c, note.get("content", "")))
            if total > 0:
                sims = self.vector_index.index.score_rows(self.vector_index._as_rows(q_emb), cand, clamp_nonneg=True)
                for r, pool in enumerate(notes_pools):
                    out[r] = [float(x) for x in sims[r, :len(pool)]]
            if dynamic:
                qn = np.asarray(q_emb.cpu() if torch is not None and isinstance(q_emb, torch.Tensor) else q_emb, dtype=np.float32)
                emb = np.asarray(self.embedding_manager.encode_queries([t for _, _, t in dynamic]), dtype=np.float32)
                for (r, c, _), e in zip(dynamic, emb):
                    sim = np.dot(qn[r], e) / (np.linalg.norm(qn[r]) * np.linalg.norm(e))
                    out[r][c] = float(max(0.0, sim))
            return out
        except Exception as e:
            logger.error(f"Error calculating vector similarities: {e}")
            return [[0.0] * len(p) for p in notes_pools]

    def search(self, queries: List[str], top_k: Optional[int] = None,
               similarity_threshold: Optional[float] = None,
               include_metadata: bool = True) -> List[List[Dict[str, Any]]]:
        if not queries:
            return []
        if not self.atomic_notes or self.vector_index.total_vectors == 0:
            logger.warning("Vector index is empty")
            return [[] for _ in queries]
        top_k = top_k or self.top_k                                                          # :199 (0 -> default)
        similarity_threshold = similarity_threshold or self.similarity_threshold             # :200
        try:
            logger.info(f"Searching for {len(queries)} queries, top_k={top_k}")
            query_embeddings = self.embedding_manager.encode_queries(queries)
            if self._size(query_embeddings) == 0:
                logger.error("Failed to generate query embeddings")
                return [[] for _ in queries]
            search_results = self.vector_index.search(query_embeddings, top_k=top_k)
            if len(queries) == 1 and isinstance(search_results, list) and \
                    len(search_results) > 0 and isinstance(search_results[0], dict):        # :222-224
                search_results = [search_results]
            final_results = []
            n_notes = len(self.atomic_notes)
            for query_idx, query in enumerate(queries):
                query_results = []
                if query_idx < len(search_results):
                    for result in search_results[query_idx]:
                        if result.get("similarity", 0) < similarity_threshold:               # :232
                            continue
                        note_index = result["index"]
                        if note_index < n_notes:
                            note = self.atomic_notes[note_index].copy()                      # :238
                            retrieval_info = {"similarity": result["similarity"], "score": result["score"],
                                              "rank": result["rank"], "query": query,
                                              "retrieval_method": "vector_search"}
                            if include_metadata:
                                note["retrieval_info"] = retrieval_info
                            else:
                                note = {"note_id": note.get("note_id"), "content": note.get("content"),
                                        "paragraph_idxs": note.get("paragraph_idxs", []),
                                        "retrieval_info": retrieval_info}
                            query_results.append(note)
                final_results.append(query_results)
            logger.info(f"Search completed: {sum(len(r) for r in final_results)} results for {len(queries)} queries")
            return final_results
        except Exception as e:
            logger.error(f"Failed to search: {e}")
            return [[] for _ in queries]

    def search_single(self, query: str, top_k: Optional[int] = None, similarity_threshold: Optional[float] = None,
                      include_metadata: bool = True) -> List[Dict[str, Any]]:
        results = self.search([query], top_k, similarity_threshold, include_metadata)
        return results[0] if results else []

    # ------------------------------------------------------------------ hybrid
    def hybrid_search(self, queries: List[str], top_k: Optional[int] = None,
                      similarity_threshold: Optional[float] = None, include_metadata: bool = True,
                      **kwargs) -> List[List[Dict[str, Any]]]:
        if not self.enable_hybrid_search or not self.hybrid_searcher or self.bm25_corpus is None:
            logger.warning("Hybrid search not available, falling back to vector search")   # :281-283
            return self.search(queries, top_k, similarity_threshold, include_metadata)
        try:
            top_k = top_k or self.top_k
            k_bm25 = int(kwargs.get("bm25_top_k", top_k))
            emb = self.embedding_manager.encode_queries(queries)
            d_scores, d_ids = self.vector_index.search_tensors(emb, top_k)
            b_scores, b_ids = self.bm25_corpus.search_batch(queries, k_bm25)
            if torch is not None and isinstance(d_scores, torch.Tensor):
                d_scores, d_ids = d_scores.cpu().numpy(), d_ids.cpu().numpy()
            fused = self.hybrid_searcher.fuse_batch(dense=(d_ids, d_scores), bm25=(b_ids, b_scores))
            out = []
            for qi, query in enumerate(queries):
                res = []
                for rank, item in enumerate(fused[qi]):
                    idx = item["note_id"]
                    if idx >= len(self.atomic_notes):
                        continue
                    note = self.atomic_notes[idx].copy()
                    info = {"similarity": item["final_similarity"], "score": item["final_similarity"], "rank": rank,
                            "query": query, "retrieval_method": "hybrid_search", "scores": item["scores"]}
                    if include_metadata:
                        note["retrieval_info"] = info
                    else:
                        note = {"note_id": note.get("note_id"), "content": note.get("content"),
                                "paragraph_idxs": note.get("paragraph_idxs", []), "retrieval_info": info}
                    res.append(note)
                out.append(res[:top_k])
            return out
        except Exception as e:
            logger.error(f"Hybrid search failed: {e}")
            logger.info("Falling back to vector search")
            return self.search(queries, top_k, similarity_threshold, include_metadata)

    def hybrid_search_single(self, query: str, top_k: Optional[int] = None,
                             similarity_threshold: Optional[float] = None, include_metadata: bool = True,
                             **kwargs) -> List[Dict[str, Any]]:
        results = self.hybrid_search([query], top_k, similarity_threshold, include_metadata, **kwargs)
        return results[0] if results else []

    # ------------------------------------------------------------------ retrieve (boost / penalty wrapper)
    def retrieve(self, query: str, top_k: Optional[int] = None, similarity_threshold: Optional[float] = None,
                 filter_fn: Optional[Callable] = None, must_have_terms: Optional[List[str]] = None,
                 boost_entities: Optional[List[str]] = None, boost_predicates: Optional[List[str]] = None,
                 topk_multiplier: float = 3.0, include_metadata: bool = True,
                 enable_guardrail: Optional[bool] = None) -> List[Dict[str, Any]]:
        if not query:
            return []
        if not self.atomic_notes or self.vector_index.total_vectors == 0:
            logger.warning("Vector index is empty")
            return []
        top_k = top_k or self.top_k
        similarity_threshold = similarity_threshold or self.similarity_threshold
        try:
            mult = topk_multiplier if topk_multiplier is not None else self.default_topk_multiplier
            expanded_top_k = int(top_k * mult)                                               # :381
            # similarity_threshold=0.0 falls back to the default through `x or default` (:200, SURVEY B)
            candidates = self.search_single(query=query, top_k=expanded_top_k, similarity_threshold=0.0,
                                            include_metadata=include_metadata)
            if not candidates:
                return []
            if filter_fn:
                kept = []
                for cand in candidates:
                    try:
                        if filter_fn(cand):
                            kept.append(cand)
                    except Exception as e:
                        logger.warning(f"Filter function failed for candidate: {e}")
                        kept.append(cand)
                candidates = kept
            enhanced = []
            for cand in candidates:
                content_lower = _note_text(cand).lower()
                original = cand.get("retrieval_info", {}).get("similarity", 0.0)
                adjusted = original
                adjustments = []
                if must_have_terms and not any(t.lower() in content_lower for t in must_have_terms):
                    adjusted *= self.must_have_terms_penalty                                 # :435
                    adjustments.append("downweighted_missing_terms")
                if boost_entities:
                    matched = [e for e in boost_entities if e.lower() in content_lower]
                    if matched:
                        adjusted *= self.entity_boost_factor                                 # :447
                        adjustments.append(f"boosted_entities_{len(matched)}")
                if boost_predicates:
                    matched = [p for p in boost_predicates if p.lower() in content_lower]
                    if matched:
                        adjusted *= self.predicate_boost_factor                              # :460
                        adjustments.append(f"boosted_predicates_{len(matched)}")
                c = cand.copy()
                if "retrieval_info" in c:
                    c["retrieval_info"] = dict(c["retrieval_info"], similarity=adjusted,
                                               original_similarity=original, adjustments=adjustments)
                enhanced.append(c)
            final = [c for c in enhanced if c.get("retrieval_info", {}).get("similarity", 0.0) >= similarity_threshold]
            final.sort(key=lambda x: x.get("retrieval_info", {}).get("similarity", 0.0), reverse=True)   # :496-499
            return final[:top_k]
        except Exception as e:
            logger.error(f"Enhanced retrieval failed: {e}")
            return self.search_single(query, top_k, similarity_threshold, include_metadata)

    # ------------------------------------------------------------------ note bookkeeping
    def add_notes(self, new_notes: List[Dict[str, Any]], rebuild_index: bool = False) -> bool:
        if not new_notes:
            return True
        try:
            if rebuild_index or not self.atomic_notes:
                return self.build_index(self.atomic_notes + new_notes, force_rebuild=True)
            start = len(self.atomic_notes)
            new_emb = self.embedding_manager.encode_atomic_notes(new_notes, include_metadata=True)
            if self._size(new_emb) == 0:
                logger.error("Failed to generate embeddings for new notes")
                return False
            new_ids = np.arange(start, start + len(new_notes), dtype=np.int64)
            if not self.vector_index.add_vectors(new_emb, new_ids):
                logger.error("Failed to add new vectors to index")
                return False
            self.atomic_notes.extend(new_notes)
            self._build_id_mappings()
            if self.note_embeddings is not None and not (torch is not None and isinstance(new_emb, torch.Tensor)):
                self.note_embeddings = np.vstack([np.asarray(self.note_embeddings), np.asarray(new_emb)])
            elif self.note_embeddings is not None:
                self.note_embeddings = torch.cat([self.note_embeddings, new_emb], dim=0)
            else:
                self.note_embeddings = new_emb
            if self.bm25_corpus is not None:
                from ..utils.bm25_search import build_bm25_corpus
                self.bm25_corpus = build_bm25_corpus(self.atomic_notes, _note_text)
            return True
        except Exception as e:
            logger.error(f"Failed to add notes: {e}")
            return False

    def remove_notes(self, note_ids: List[str]) -> bool:
        if not note_ids:
            return True
        try:
            idxs = sorted({self.note_id_to_index[n] for n in note_ids if n in self.note_id_to_index}, reverse=True)
            if not idxs:
                logger.warning("No matching notes found to remove")
                return True
            for idx in idxs:
                if idx < len(self.atomic_notes):
                    del self.atomic_notes[idx]
            return self.build_index(self.atomic_notes, force_rebuild=True)                   # :591
        except Exception as e:
            logger.error(f"Failed to remove notes: {e}")
            return False

    def update_note(self, note_id: str, updated_note: Dict[str, Any]) -> bool:
        try:
            if note_id not in self.note_id_to_index:
                logger.warning(f"Note {note_id} not found")
                return False
            self.atomic_notes[self.note_id_to_index[note_id]] = updated_note
            return self.build_index(self.atomic_notes, force_rebuild=True)                   # :624
        except Exception as e:
            logger.error(f"Failed to update note {note_id}: {e}")
            return False

    def get_note_by_id(self, note_id: str) -> Optional[Dict[str, Any]]:
        idx = self.note_id_to_index.get(note_id)
        if idx is not None and idx < len(self.atomic_notes):
            return self.atomic_notes[idx]
        return None

    def get_notes_by_ids(self, note_ids: List[str]) -> List[Dict[str, Any]]:
        return [n for n in (self.get_note_by_id(i) for i in note_ids) if n]

    def get_similar_notes(self, note_id: str, top_k: int = 10, exclude_self: bool = True) -> List[Dict[str, Any]]:
        note = self.get_note_by_id(note_id)
        if not note or not note.get("content", ""):
            return []
        results = self.search_single(note["content"], top_k=top_k + (1 if exclude_self else 0))
        if exclude_self:
            results = [r for r in results if r.get("note_id") != note_id]
        return results[:top_k]

    def _build_id_mappings(self):
        self.note_id_to_index = {}
        self.index_to_note_id = {}
        for idx, note in enumerate(self.atomic_notes):
            note_id = note.get("note_id")
            if note_id:
                self.note_id_to_index[note_id] = idx
                self.index_to_note_id[idx] = note_id

    # ------------------------------------------------------------------ persistence (:680-749)
    def _can_load_existing_index(self, atomic_notes: List[Dict[str, Any]]) -> bool:
        try:
            index_files = [f for f in os.listdir(self.data_dir) if f.endswith(".faiss")]
            data_file = os.path.join(self.data_dir, "atomic_notes.json")
            if not index_files or not os.path.exists(data_file):
                return False
            with open(data_file, "r", encoding="utf-8") as f:
                existing = json.load(f)
            if len(existing) != len(atomic_notes):
                return False
            if existing and atomic_notes and existing[0].get("note_id") != atomic_notes[0].get("note_id"):
                return False
            old_dir = self.vector_index.index_dir
            self.vector_index.index_dir = self.data_dir
            ok = self.vector_index.load_index(index_files[0])
            self.vector_index.index_dir = old_dir
            if ok:
                self.atomic_notes = existing
                self._build_id_mappings()
                emb_file = os.path.join(self.data_dir, "note_embeddings.npz")
                if os.path.exists(emb_file):
                    self.note_embeddings = np.load(emb_file)["embeddings"]
                return True
        except Exception as e:
            logger.warning(f"Failed to load existing index: {e}")
        return False

    def _save_index_data(self):
        try:
            old_dir = self.vector_index.index_dir
            self.vector_index.index_dir = self.data_dir
            self.vector_index.save_index()
            self.vector_index.index_dir = old_dir
            with open(os.path.join(self.data_dir, "atomic_notes.json"), "w", encoding="utf-8") as f:
                json.dump(self.atomic_notes, f, ensure_ascii=False)
            if self.note_embeddings is not None:
                emb = self.note_embeddings
                if torch is not None and isinstance(emb, torch.Tensor):
                    emb = emb.float().cpu().numpy()
                np.savez_compressed(os.path.join(self.data_dir, "note_embeddings.npz"), embeddings=emb)
            with open(os.path.join(self.data_dir, "id_mappings.json"), "w", encoding="utf-8") as f:
                json.dump({"note_id_to_index": self.note_id_to_index,
                           "index_to_note_id": {str(k): v for k, v in self.index_to_note_id.items()}}, f,
                          ensure_ascii=False)
            logger.info("Index data saved successfully")
        except Exception as e:
            logger.error(f"Failed to save index data: {e}")

    def get_retrieval_stats(self) -> Dict[str, Any]:
        return {"total_notes": len(self.atomic_notes),
                "embedding_dim": self.embedding_manager.embedding_dim,
                "model_name": getattr(self.embedding_manager, "model_name", None),
                "index_stats": self.vector_index.get_index_stats(),
                "top_k": self.top_k, "similarity_threshold": self.similarity_threshold}

    # ------------------------------------------------------------------ TF-IDF "BM25 fallback" (:924-1032)
    def _build_bm25_index(self, atomic_notes: List[Dict[str, Any]]) -> None:
        """sklearn's TfidfVectorizer over the pre-processed contents, exactly as the reference configures it
        (:941-946); the L2-normalised matrix goes to the device (``engine.TfidfIndex``)."""
        try:
            from sklearn.feature_extraction.text import TfidfVectorizer
            from sklearn.preprocessing import normalize
            texts = [self._preprocess_text(note.get("content", "")) if note.get("content", "") else "" for note in atomic_notes]
            self.processed_texts = texts
            self.tfidf_vectorizer = TfidfVectorizer(lowercase=True, stop_words="english", max_features=10000, ngram_range=(1, 2))
            if texts:
                self.tfidf_matrix = self.tfidf_vectorizer.fit_transform(texts)
                if self._tfidf_index is not None:
                    self._tfidf_index.close()
                # cosine_similarity normalises its operands again (sklearn.metrics.pairwise): do it once here
                self._tfidf_index = engine.TfidfIndex(normalize(self.tfidf_matrix, copy=True))
                logger.info(f"BM25 index built with {len(texts)} documents")
            else:
                logger.warning("No valid texts found for BM25 indexing")
        except Exception as e:
            logger.error(f"Failed to build BM25 index: {e}")
            self.bm25_enabled = False

    @staticmethod
    def _preprocess_text(text: str) -> str:
        """:958-964"""
        import re
        text = re.sub(r"[^a-zA-Z0-9\s]", " ", text)
        text = re.sub(r"\s+", " ", text)
        return text.strip().lower()

    def _bm25_search_batch(self, queries: List[str], top_k: int = 20) -> List[List[Dict[str, Any]]]:
        """Many fallback queries through one device call (the reference answers them one by one, :966)."""
        if not self.bm25_enabled or not self.tfidf_vectorizer or self._tfidf_index is None:
            return [[] for _ in queries]
        try:
            from sklearn.preprocessing import normalize
            qm = normalize(self.tfidf_vectorizer.transform([self._preprocess_text(q) for q in queries]), copy=True)
            scores, ids = self._tfidf_index.search(qm, int(top_k))
            out = []
            for query, srow, irow in zip(queries, scores.tolist(), ids.tolist()):
                results = []
                for sim, idx in zip(srow, irow):
                    if 0 <= idx < len(self.atomic_notes) and sim > 0:              # :985
                        note = self.atomic_notes[idx].copy()
                        note["retrieval_info"] = {"similarity": float(sim), "score": float(sim), "rank": len(results) + 1,
                                                  "query": query, "retrieval_method": "bm25_fallback"}
                        results.append(note)
                logger.info(f"BM25 fallback search returned {len(results)} results")
                out.append(results)
            return out
        except Exception as e:
            logger.error(f"BM25 search failed: {e}")
            return [[] for _ in queries]

    def _bm25_search(self, query: str, top_k: int = 20) -> List[Dict[str, Any]]:
        """:966-1002"""
        return self._bm25_search_batch([query], top_k)[0]

    def search_with_namespace_fallback(self, queries: List[str], dataset: str, qid: str, top_k: Optional[int] = None,
                                       similarity_threshold: Optional[float] = None, include_metadata: bool = True,
                                       namespace_filter: Optional[Callable] = None) -> List[List[Dict[str, Any]]]:
        """:1004-1032.  ``namespace_filter(notes, dataset, qid)`` defaults to the reference's
        ``utils.dataset_guard.filter_notes_by_namespace`` (imported at call time, as the reference does); the
        queries whose vector results are all filtered away share one batched fallback call."""
        if namespace_filter is None:
            from utils.dataset_guard import filter_notes_by_namespace as namespace_filter
        vector_results = self.search(queries, top_k, similarity_threshold, include_metadata)
        final = [namespace_filter(res, dataset, qid) for res in vector_results]
        need = [i for i, res in enumerate(final) if not res] if self.bm25_enabled else []
        if need:
            fallback = self._bm25_search_batch([queries[i] for i in need], top_k or self.top_k)
            for i, res in zip(need, fallback):
                logger.info(f"No namespace matches for query {i + 1}, trying BM25 fallback")
                kept = namespace_filter(res, dataset, qid)
                if kept:
                    logger.info(f"BM25 fallback found {len(kept)} namespace matches")
                    final[i] = kept
                else:
                    logger.warning(f"No results found for query '{queries[i]}' in namespace {dataset}/{qid}")
        return final

    def cleanup(self):
        try:
            if getattr(self, "_tfidf_index", None) is not None:
                self._tfidf_index.close()
                self._tfidf_index = None
            self.vector_index.cleanup()
            if self.bm25_corpus is not None and hasattr(self.bm25_corpus, "close"):
                self.bm25_corpus.close()
            self.bm25_corpus = None
        except Exception as e:
            logger.error(f"Cleanup failed: {e}")
