"""Oracle: BM25 scoring (TEST INFRASTRUCTURE ONLY).

* ``SimpleBM25Oracle`` restates ``utils/bm25_search.py:16-63`` (PINNED against the
  reference file loaded by path, see tests/golden/).
* ``BM25OkapiOracle`` restates ``rank_bm25.BM25Okapi`` (rank_bm25 0.2.x; the
  reference does not pin it and it is not in requirements.txt -- optional import
  at ``utils/bm25_search.py:7-13``, constructed with library defaults k1=1.5,
  b=0.75, epsilon=0.25 at ``:273``).  The library is not installed here; the
  restatement follows its published algorithm and is PINNED ONLY to the one
  library-held vector available offline, the known answer of rank_bm25's README
  (corpus of three sentences, query "windy London" -> [0., 0.93729472, 0.];
  tests/test_oracle_golden.py::test_okapi_oracle_reproduces_rank_bm25_readme_vector).
  Beyond that vector: PARITY UNPINNED.
* ``bm25_scores`` restates the wrapper ``utils/bm25_search.py:286-340``.

All arithmetic is IEEE float64 in the reference's operation order; that order is
what the CUDA path has to reproduce bit-for-bit.
"""
from __future__ import annotations

import math
import re
from collections import Counter
from typing import Any, Callable, Dict, List, Sequence

import numpy as np


def tokenize_text(text: str) -> List[str]:
    """utils/bm25_search.py:237-241."""
    return re.findall(r"\b\w+\b", text.lower())


class SimpleBM25Oracle:
    """utils/bm25_search.py:16-63 (Lucene-style idf, ``b * (dl / avgdl)``)."""

    def __init__(self, corpus: List[List[str]], k1: float = 1.5, b: float = 0.75):
        self.corpus = corpus
        self.k1 = k1
        self.b = b
        self.doc_len = [len(doc) for doc in corpus]
        self.avgdl = sum(self.doc_len) / len(self.doc_len) if self.doc_len else 0
        self.doc_freqs = [Counter(doc) for doc in corpus]
        self.doc_count = len(corpus)
        df: Dict[str, int] = {}
        for freq in self.doc_freqs:
            for term in freq:
                df[term] = df.get(term, 0) + 1
        # :39-41 (the reference recounts df per term in O(V*N); same value)
        self.idf = {t: math.log((self.doc_count - n + 0.5) / (n + 0.5) + 1.0) for t, n in df.items()}

    def get_scores(self, query: List[str]) -> List[float]:
        scores = []
        for i, doc_freq in enumerate(self.doc_freqs):          # :47
            score = 0.0
            doc_len = self.doc_len[i]
            for term in query:                                  # :51
                if term in doc_freq:
                    tf = doc_freq[term]
                    idf = self.idf.get(term, 0)
                    numerator = tf * (self.k1 + 1)              # :57
                    denominator = tf + self.k1 * (1 - self.b + self.b * (doc_len / self.avgdl))  # :58
                    score += idf * (numerator / denominator)    # :59
            scores.append(score)
        return scores


class BM25OkapiOracle:
    """rank_bm25.BM25Okapi restated line for line (PARITY UNPINNED).

    __init__/_initialize: df counted per document, ``avgdl = total_tokens / N``
    (ZeroDivisionError on an empty corpus, as in the library).
    _calc_idf: ``idf = ln(N - n + 0.5) - ln(n + 0.5)``; ``average_idf`` is the sum
    over the vocabulary IN FIRST-APPEARANCE ORDER divided by |V|; negative idfs
    become ``epsilon * average_idf``.
    get_scores: for each query token in order (duplicates repeat)
    ``score += idf * (tf*(k1+1) / (tf + k1*(1 - b + b*dl/avgdl)))`` over all docs.
    """

    def __init__(self, corpus: List[List[str]], k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25):
        self.k1 = k1
        self.b = b
        self.epsilon = epsilon
        self.corpus_size = 0
        self.avgdl = 0
        self.doc_freqs: List[Dict[str, int]] = []
        self.idf: Dict[str, float] = {}
        self.doc_len: List[int] = []
        nd: Dict[str, int] = {}
        num_doc = 0
        for document in corpus:
            self.doc_len.append(len(document))
            num_doc += len(document)
            frequencies: Dict[str, int] = {}
            for word in document:
                if word not in frequencies:
                    frequencies[word] = 0
                frequencies[word] += 1
            self.doc_freqs.append(frequencies)
            for word in frequencies:
                nd[word] = nd.get(word, 0) + 1
            self.corpus_size += 1
        self.avgdl = num_doc / self.corpus_size
        idf_sum = 0
        negative_idfs = []
        for word, freq in nd.items():
            idf = math.log(self.corpus_size - freq + 0.5) - math.log(freq + 0.5)
            self.idf[word] = idf
            idf_sum += idf
            if idf < 0:
                negative_idfs.append(word)
        self.average_idf = idf_sum / len(self.idf)
        eps = self.epsilon * self.average_idf
        for word in negative_idfs:
            self.idf[word] = eps

    def get_scores(self, query: Sequence[str]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:
            q_freq = np.array([(doc.get(q) or 0) for doc in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (q_freq * (self.k1 + 1) /
                                               (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
        return score


def build_bm25_corpus(notes: List[Dict[str, Any]], text_fn: Callable[[Dict[str, Any]], str],
                      backend: str = "okapi"):
    """utils/bm25_search.py:244-283.  ``backend='okapi'`` is what runs where
    rank_bm25 is installed, ``'simple'`` is the in-repo fallback (:281)."""
    tokenized = []
    for note in notes:
        try:
            text = text_fn(note)
            tokenized.append(tokenize_text(text) if text else [])
        except Exception:
            tokenized.append([])
    if backend == "okapi":
        try:
            return BM25OkapiOracle(tokenized)
        except Exception:
            pass                                                # :276-283 fall back
    return SimpleBM25Oracle(tokenized)


def bm25_scores(corpus: Any, docs: Sequence[Any], query: str) -> List[float]:
    """utils/bm25_search.py:286-340: tokenise, score, pad/truncate, max-normalise."""
    try:
        query_tokens = tokenize_text(query)
        if not query_tokens:
            return [0.0] * len(docs)
        scores = corpus.get_scores(query_tokens)
        if hasattr(scores, "tolist"):
            scores = scores.tolist()
        if len(scores) != len(docs):
            if len(scores) < len(docs):
                scores.extend([0.0] * (len(docs) - len(scores)))
            else:
                scores = scores[:len(docs)]
        if scores:
            max_score = max(scores)
            if max_score > 0:
                scores = [s / max_score for s in scores]
        return scores
    except Exception:
        return [0.0] * len(docs)


# ---------------------------------------------------------------------------
# Array (token-id) form used at sizes where Python dict-of-str corpora are too
# slow to build.  Bit-identical to the classes above (tests assert it).
# ---------------------------------------------------------------------------

class ArrayBM25Oracle:
    """Same arithmetic as BM25OkapiOracle / SimpleBM25Oracle over integer token ids.

    ``doc_tokens``: list/array of int arrays (one per doc).  Term ids may be any
    non-negative ints; the vocabulary order used for ``average_idf`` is first
    appearance, exactly like the dict insertion order in rank_bm25.
    """

    def __init__(self, doc_ptr: np.ndarray, doc_tok: np.ndarray, variant: str = "okapi",
                 k1: float = 1.5, b: float = 0.75, epsilon: float = 0.25):
        self.variant = variant
        self.k1, self.b, self.epsilon = k1, b, epsilon
        doc_ptr = np.asarray(doc_ptr, dtype=np.int64)
        doc_tok = np.asarray(doc_tok, dtype=np.int64)
        n = len(doc_ptr) - 1
        self.n = n
        self.doc_len = np.diff(doc_ptr)
        total = int(self.doc_len.sum())
        if variant == "okapi":
            self.avgdl = total / n                    # ZeroDivisionError when n == 0
        else:
            self.avgdl = total / n if n else 0
        doc_of = np.repeat(np.arange(n, dtype=np.int64), self.doc_len)
        # (term, doc) -> tf
        key = doc_tok * n + doc_of
        uniq, first_pos, tf = np.unique(key, return_index=True, return_counts=True)
        terms = uniq // n
        docs = uniq % n
        self.post_term, self.post_doc, self.post_tf = terms, docs, tf
        vocab, term_start, df = np.unique(terms, return_index=True, return_counts=True)
        self.vocab = vocab
        self.term_start = term_start
        self.df = df
        self.term_slot = {int(t): i for i, t in enumerate(vocab)}
        # first-appearance order of the vocabulary (dict insertion order in rank_bm25)
        first_tok_pos = np.full(len(vocab), total, dtype=np.int64)
        slot_of_posting = np.searchsorted(vocab, terms)
        np.minimum.at(first_tok_pos, slot_of_posting, first_pos)
        self.appearance_order = np.argsort(first_tok_pos, kind="stable")
        self.idf = np.zeros(len(vocab), dtype=np.float64)
        if variant == "okapi":
            idf_sum = 0
            neg = []
            for slot in self.appearance_order.tolist():
                f = int(df[slot])
                v = math.log(n - f + 0.5) - math.log(f + 0.5)
                self.idf[slot] = v
                idf_sum += v
                if v < 0:
                    neg.append(slot)
            self.average_idf = idf_sum / len(vocab)   # ZeroDivisionError when no tokens
            eps = self.epsilon * self.average_idf
            for slot in neg:
                self.idf[slot] = eps
        else:
            for slot in range(len(vocab)):
                f = int(df[slot])
                self.idf[slot] = math.log((n - f + 0.5) / (f + 0.5) + 1.0)

    def impacts(self, slot: int):
        s, e = int(self.term_start[slot]), int(self.term_start[slot] + self.df[slot])
        docs = self.post_doc[s:e]
        tf = self.post_tf[s:e]
        dl = self.doc_len[docs]
        k1, b = self.k1, self.b
        if self.variant == "okapi":
            g = tf * (k1 + 1) / (tf + k1 * (1 - b + b * dl / self.avgdl))
        else:
            g = (tf * (k1 + 1)) / (tf + k1 * (1 - b + b * (dl / self.avgdl)))
        return docs, self.idf[slot] * g

    def get_scores(self, query_ids: Sequence[int]) -> np.ndarray:
        score = np.zeros(self.n)
        for t in query_ids:
            slot = self.term_slot.get(int(t))
            if slot is None:
                continue
            idf = self.idf[slot]
            if self.variant == "okapi" and not idf:
                continue                               # (idf.get(q) or 0) * ... adds +0.0
            docs, imp = self.impacts(slot)
            score[docs] += imp                         # docs are unique within a term
        return score


def topk_positive(scores: np.ndarray, k: int):
    """BM25 top-k as the GPU API defines it: descending score,
    ascending doc id on ties, docs with score <= 0 excluded, -1 padding."""
    scores = np.asarray(scores, dtype=np.float64)
    cand = np.nonzero(scores > 0)[0]
    order = np.lexsort((cand, -scores[cand]))[:k]
    sel = cand[order]
    ids = np.full(k, -1, dtype=np.int64)
    out = np.zeros(k, dtype=np.float64)
    ids[:len(sel)] = sel
    out[:len(sel)] = scores[sel]
    return out, ids
