"""BM25 over the B200 inverted-index kernel -- drop-in for reference
utils/bm25_search.py (tokenize_text :237-241, build_bm25_corpus :244-283,
bm25_scores :286-340; SimpleBM25 :16-63 and rank_bm25.BM25Okapi are replaced by
``GpuBM25``).

Host side here: tokenisation, the str -> term-id dictionary and the idf table
(``math.log``, bit-identical to the reference).  Every score is computed on the
device in fp64 in the reference's operation order (csrc/bm25.cu); there is no CPU
scorer behind this module.
"""
from __future__ import annotations

import re
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np

from .._log import logger
from ..config import config
from .. import bm25_build as bb
from .. import engine
from .._native import MAX_QUERY_TERMS

RANK_BM25_AVAILABLE = False     # kept for import compatibility (bm25_search.py:7-13); the GPU back-end is used

_TOKEN_RE = re.compile(r"\b\w+\b")


def tokenize_text(text: str) -> List[str]:
    """Lower-case word tokens (reference :237-241)."""
    return _TOKEN_RE.findall(text.lower())


class GpuBM25:
    """Corpus object returned by ``build_bm25_corpus``.

    variant='okapi' reproduces rank_bm25.BM25Okapi (k1=1.5, b=0.75, epsilon=0.25: the
    library defaults, because the reference passes none at :273); variant='simple'
    reproduces the in-repo SimpleBM25 (:16-63)."""

    def __init__(self, tokenized_corpus: List[List[str]], variant: str = "okapi", k1: float = bb.K1_DEFAULT,
                 b: float = bb.B_DEFAULT, epsilon: float = bb.EPS_DEFAULT):
        self.variant = variant
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.vocab = bb.Vocabulary()
        doc_ptr, doc_tok = self.vocab.encode_docs(tokenized_corpus)
        arrays = bb.build_arrays(doc_ptr, doc_tok, len(self.vocab), variant=variant, k1=k1, b=b, epsilon=epsilon)
        self._finish(arrays)

    @classmethod
    def from_token_ids(cls, doc_ptr: np.ndarray, doc_tok: np.ndarray, n_terms: int, variant: str = "okapi",
                       k1: float = bb.K1_DEFAULT, b: float = bb.B_DEFAULT, epsilon: float = bb.EPS_DEFAULT,
                       stats: Optional[bb.BM25Stats] = None) -> "GpuBM25":
        """Corpus given as integer term ids (synthetic corpora, shards with global ``stats``)."""
        self = cls.__new__(cls)
        self.variant = variant
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.vocab = None
        self._finish(bb.build_arrays(doc_ptr, doc_tok, n_terms, variant=variant, k1=k1, b=b, epsilon=epsilon,
                                     stats=stats))
        return self

    def _finish(self, arrays: bb.BM25Arrays) -> None:
        st = arrays.stats
        self.stats = st
        self.corpus_size = self.doc_count = arrays.n_docs_local
        self.avgdl = st.avgdl
        self.doc_len = arrays.doc_len.tolist() if arrays.n_docs_local <= 1_000_000 else arrays.doc_len
        self.average_idf = st.average_idf
        self.index = engine.BM25Index(arrays.n_docs_local, arrays.term_ptr, arrays.post_doc, arrays.post_gcode,
                                      st.idf, arrays.gtab)
        self.n_g = len(arrays.gtab)

    @property
    def idf(self) -> Dict[str, float]:
        if self.vocab is None:
            return {}
        idf = self.stats.idf
        return {w: float(idf[t]) for w, t in self.vocab.ids.items()}

    # -- encoding of queries into padded term-id matrices
    def _term_ids(self, toks: Sequence[Any]) -> List[int]:
        """Known term ids of a tokenised query, in token order.  Tokens outside the vocabulary add +0.0 in the
        reference (``idf.get(q) or 0``): leaving them out changes nothing and keeps the id matrix narrow."""
        ids = self.vocab.encode_query(toks) if self.vocab is not None else [int(t) for t in toks]
        return [t for t in ids if t >= 0]

    def encode_queries(self, token_lists: Sequence[Sequence[Any]]) -> np.ndarray:
        rows = [self._term_ids(toks) if toks else [] for toks in token_lists]
        width = max(1, max((len(r) for r in rows), default=1))
        if width > MAX_QUERY_TERMS:
            raise ValueError(f"queries with more than {MAX_QUERY_TERMS} known tokens need scores_long()")
        out = np.full((len(rows), width), -1, dtype=np.int32)
        for i, ids in enumerate(rows):
            out[i, :len(ids)] = ids
        return out

    def scores_long(self, query_tokens: Sequence[Any], n_out: Optional[int] = None, normalise: bool = False) -> np.ndarray:
        """Scores of every doc for ONE query of any length (the reference has no limit): the known tokens are scored
        ``MAX_QUERY_TERMS`` at a time, every chunk continuing the sums of the previous one on the device, so the
        fp64 summation order is the reference's token order throughout."""
        ids = self._term_ids(list(query_tokens))
        chunks = [ids[i:i + MAX_QUERY_TERMS] for i in range(0, max(1, len(ids)), MAX_QUERY_TERMS)]
        acc = None
        for j, ch in enumerate(chunks):
            qt = np.asarray([ch if ch else [-1]], dtype=np.int32)
            acc = self.index.scores_device(qt, n_out=n_out, normalise=normalise and j == len(chunks) - 1, continue_from=acc)
        return acc[0].cpu().numpy()

    # -- reference-shaped scoring
    def get_scores(self, query: Sequence[Any]) -> np.ndarray:
        """Raw scores of every doc for one tokenised query (BM25Okapi.get_scores contract:
        float64 ndarray of length N)."""
        query = list(query)
        if len(query) > MAX_QUERY_TERMS:
            return self.scores_long(query)
        return self.index.scores(self.encode_queries([query]))[0]

    # -- batched GPU paths (additions)
    def scores_batch(self, queries: Sequence[str], n_out: Optional[int] = None, normalise: bool = True) -> np.ndarray:
        return self.index.scores(self.encode_queries([tokenize_text(q) for q in queries]), n_out=n_out,
                                 normalise=normalise)

    def search_batch(self, queries: Sequence[str], k: int, id_base: int = 0):
        """Top-k per query: (scores fp64 [nq,k], ids int64 [nq,k]), score desc / doc asc, zero scores excluded."""
        return self.index.search(self.encode_queries([tokenize_text(q) for q in queries]), k, id_base=id_base)

    def search_terms(self, q_terms, k: int, id_base: int = 0):
        return self.index.search(q_terms, k, id_base=id_base)

    def close(self):
        self.index.close()


class FieldWeightedBM25:
    """Reference :66-145: per-field (title, entities, content) BM25 with SimpleBM25 arithmetic and
    per-field idf / avgdl, combined as ``total += weight * field_score`` in field order.  Each field is
    its own device index; the combination and the max-normalisation run on the device (csrc/bm25.cu)."""

    def __init__(self, corpus: List[Dict[str, List[str]]], field_weights: Dict[str, float] = None, k1: float = 1.5,
                 b: float = 0.75):
        self.corpus = corpus
        self.field_weights = field_weights or {"title": 2.0, "entities": 1.5, "content": 1.0}
        self.k1, self.b = k1, b
        self.doc_count = len(corpus)
        self.fields = {f: GpuBM25([doc.get(f, []) for doc in corpus], variant="simple", k1=k1, b=b)
                       for f in self.field_weights}
        self.field_stats = {f: {"doc_lens": c.doc_len, "avgdl": c.avgdl} for f, c in self.fields.items()}
        self.field_idf = {f: c.idf for f, c in self.fields.items()}

    def _scores_device(self, query_tokens: Sequence[str], normalise: bool, n_out: Optional[int] = None):
        per_field = []
        for c in self.fields.values():
            ids = c._term_ids(list(query_tokens))
            acc = None
            for i in range(0, max(1, len(ids)), MAX_QUERY_TERMS):       # any query length: chunks continue the same sums
                qt = np.asarray([ids[i:i + MAX_QUERY_TERMS] or [-1]], dtype=np.int32)
                acc = c.index.scores_device(qt, n_out=n_out, continue_from=acc)
            per_field.append(acc)
        return engine.combine_fields(per_field, list(self.field_weights.values()), normalise)

    def get_scores(self, query: List[str]) -> List[float]:
        if self.doc_count == 0:
            return []
        return self._scores_device(query, False)[0].cpu().numpy().tolist()

    def close(self):
        for c in self.fields.values():
            c.close()


def build_field_weighted_bm25_corpus(notes: List[Dict[str, Any]], field_weights: Dict[str, float] = None) -> FieldWeightedBM25:
    """Reference :148-190: title / entities (joined list) / content, each tokenised with tokenize_text."""
    if field_weights is None:
        field_weights = {"title": 2.0, "entities": 1.5, "content": 1.0}
    field_corpus = []
    for note in notes:
        entities = note.get("entities", []) or []
        entities_text = " ".join(entities) if isinstance(entities, list) else str(entities)
        field_corpus.append({"title": tokenize_text(note.get("title", "") or ""),
                             "entities": tokenize_text(entities_text),
                             "content": tokenize_text(note.get("content", "") or "")})
    return FieldWeightedBM25(field_corpus, field_weights)


def field_weighted_bm25_scores(corpus: FieldWeightedBM25, docs: List[Dict[str, Any]], query: str) -> List[float]:
    """Reference :193-234: same wrapper as bm25_scores (pad / truncate to len(docs), max-normalise)."""
    try:
        query_tokens = tokenize_text(query)
        if not query_tokens:
            logger.warning("Empty query tokens, returning zero scores")
            return [0.0] * len(docs)
        if len(docs) == 0:
            return []
        return corpus._scores_device(query_tokens, True, n_out=len(docs))[0].cpu().numpy().tolist()
    except Exception as e:
        logger.error(f"Error calculating field-weighted BM25 scores: {e}")
        return [0.0] * len(docs)


def build_bm25_corpus(notes: List[Dict[str, Any]], text_fn: Callable[[Dict[str, Any]], str]) -> Any:
    """Reference :244-283: tokenise every note (``[]`` on empty text or error), then
    build the corpus object.  Okapi first (what runs wherever rank_bm25 is installed);
    its constructor raises on degenerate corpora exactly where the library does, and
    the SimpleBM25 arithmetic takes over as in the reference (:276-283)."""
    tokenized_corpus: List[List[str]] = []
    for note in notes:
        try:
            text = text_fn(note)
            tokenized_corpus.append(tokenize_text(text) if text else [])
        except Exception as e:
            logger.warning(f"Error extracting text from note: {e}")
            tokenized_corpus.append([])
    backend = config.get("b200.bm25_backend", "okapi")
    if backend == "okapi":
        try:
            corpus = GpuBM25(tokenized_corpus, variant="okapi")
            logger.debug(f"Built BM25 corpus with {len(tokenized_corpus)} documents (okapi arithmetic, B200)")
            return corpus
        except ZeroDivisionError as e:
            logger.error(f"Error building BM25 corpus with okapi arithmetic: {e}")
    corpus = GpuBM25(tokenized_corpus, variant="simple")
    logger.debug(f"Built BM25 corpus with {len(tokenized_corpus)} documents (simple arithmetic, B200)")
    return corpus


def bm25_scores(corpus: Any, docs: List[Dict[str, Any]], query: str) -> List[float]:
    """Reference :286-340: scores of the WHOLE corpus for ``query``, padded / truncated to
    ``len(docs)`` (:320-327), divided by their max when it is positive (:329-333); zeros
    on an empty query (:302-304) or any error (:338-340)."""
    try:
        query_tokens = tokenize_text(query)
        if not query_tokens:
            logger.warning("Empty query tokens, returning zero scores")
            return [0.0] * len(docs)
        if isinstance(corpus, GpuBM25):
            if len(docs) == 0:
                return []
            if len(query_tokens) > MAX_QUERY_TERMS:
                return corpus.scores_long(query_tokens, n_out=len(docs), normalise=True).tolist()
            return corpus.index.scores(corpus.encode_queries([query_tokens]), n_out=len(docs),
                                       normalise=True)[0].tolist()
        if isinstance(corpus, FieldWeightedBM25):
            return field_weighted_bm25_scores(corpus, docs, query)
        if hasattr(corpus, "get_scores"):
            # any other corpus object with the reference's duck type (:307-314): its own scorer, then the same
            # pad / truncate / max-normalise steps (:320-333)
            scores = corpus.get_scores(query_tokens)
            scores = scores.tolist() if hasattr(scores, "tolist") else list(scores)
            if len(scores) < len(docs):
                scores = scores + [0.0] * (len(docs) - len(scores))
            elif len(scores) > len(docs):
                scores = scores[:len(docs)]
            max_score = max(scores) if scores else 0.0
            if max_score > 0:
                scores = [s / max_score for s in scores]
            return scores
        logger.error("Invalid corpus object")
        return [0.0] * len(docs)
    except Exception as e:
        logger.error(f"Error calculating BM25 scores: {e}")
        return [0.0] * len(docs)
