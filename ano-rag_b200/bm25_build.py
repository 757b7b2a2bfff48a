"""Host-side construction of the BM25 inverted index (build time, not the hot path).

Follows what the reference does on the host at index build
(utils/bm25_search.py:244-283 -> rank_bm25.BM25Okapi.__init__/_calc_idf, or
SimpleBM25.__init__ :19-41): tokenised docs -> df -> idf with ``math.log`` (libm,
so the values are bit-identical to the reference's; a device log would not be) ->
per-(tf, doc_len) factor table.  The arrays go to the device through
``engine.BM25Index`` (C ABI ``arb_bm25_create``); scoring happens only there.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

K1_DEFAULT, B_DEFAULT, EPS_DEFAULT = 1.5, 0.75, 0.25   # rank_bm25 / SimpleBM25 defaults (bm25_search.py:19,273)


@dataclass
class BM25Stats:
    """Corpus-global statistics; identical on every shard."""
    variant: str                 # 'okapi' | 'simple'
    n_docs: int
    avgdl: float
    df: np.ndarray               # int64 [V]
    idf: np.ndarray              # float64 [V]
    k1: float = K1_DEFAULT
    b: float = B_DEFAULT
    epsilon: float = EPS_DEFAULT
    average_idf: Optional[float] = None


def term_doc_counts(doc_ptr: np.ndarray, doc_tok: np.ndarray, n_terms: int):
    """(term, doc) -> tf as a term-major, doc-sorted CSR.  Returns term_ptr, post_doc, post_tf."""
    doc_ptr = np.asarray(doc_ptr, dtype=np.int64)
    doc_tok = np.asarray(doc_tok, dtype=np.int64)
    n = len(doc_ptr) - 1
    if n == 0 or len(doc_tok) == 0:
        return np.zeros(n_terms + 1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int64)
    doc_of = np.repeat(np.arange(n, dtype=np.int64), np.diff(doc_ptr))
    key = doc_tok * n + doc_of
    key.sort(kind="stable")
    new = np.empty(len(key), dtype=bool)
    new[0] = True
    np.not_equal(key[1:], key[:-1], out=new[1:])
    starts = np.flatnonzero(new)
    uniq = key[starts]
    tf = np.diff(np.append(starts, len(key)))
    terms = uniq // n
    docs = (uniq % n).astype(np.int32)
    term_ptr = np.zeros(n_terms + 1, dtype=np.int64)
    np.cumsum(np.bincount(terms, minlength=n_terms), out=term_ptr[1:])
    return term_ptr, docs, tf.astype(np.int64)


def first_appearance_order(doc_tok: np.ndarray, n_terms: int) -> np.ndarray:
    """Term ids ordered by first occurrence in the token stream (dict insertion order
    of rank_bm25's ``nd``); terms that never occur are left out."""
    doc_tok = np.asarray(doc_tok, dtype=np.int64)
    first = np.full(n_terms, len(doc_tok), dtype=np.int64)
    np.minimum.at(first, doc_tok, np.arange(len(doc_tok), dtype=np.int64))
    present = np.flatnonzero(first < len(doc_tok))
    return present[np.argsort(first[present], kind="stable")]


def compute_stats(variant: str, n_docs: int, total_tokens: int, df: np.ndarray, appearance: Sequence[int],
                  k1: float = K1_DEFAULT, b: float = B_DEFAULT, epsilon: float = EPS_DEFAULT) -> BM25Stats:
    """idf exactly as the reference computes it (Python floats, libm log)."""
    df = np.asarray(df, dtype=np.int64)
    idf = np.zeros(len(df), dtype=np.float64)
    average_idf = None
    if variant == "okapi":
        avgdl = total_tokens / n_docs                       # ZeroDivisionError on an empty corpus, like rank_bm25
        idf_sum = 0
        negative = []
        for t in appearance:
            f = int(df[t])
            v = math.log(n_docs - f + 0.5) - math.log(f + 0.5)
            idf[t] = v
            idf_sum += v
            if v < 0:
                negative.append(t)
        average_idf = idf_sum / len(appearance)             # ZeroDivisionError when there are no tokens
        eps = epsilon * average_idf
        for t in negative:
            idf[t] = eps
    elif variant == "simple":
        avgdl = total_tokens / n_docs if n_docs else 0      # bm25_search.py:23
        for t in appearance:
            f = int(df[t])
            idf[t] = math.log((n_docs - f + 0.5) / (f + 0.5) + 1.0)   # :41
    else:
        raise ValueError(f"unknown BM25 variant {variant!r}")
    return BM25Stats(variant, int(n_docs), avgdl, df, idf, k1, b, epsilon, average_idf)


def factor_table(stats: BM25Stats, post_tf: np.ndarray, post_dl: np.ndarray):
    """Distinct (tf, dl) pairs -> table of the reference's saturation factor, plus the
    per-posting table index.  The numpy expressions are the reference's own
    (rank_bm25 get_scores; SimpleBM25.get_scores bm25_search.py:57-59)."""
    post_tf = np.asarray(post_tf, dtype=np.int64)
    post_dl = np.asarray(post_dl, dtype=np.int64)
    if len(post_tf) == 0:
        return np.zeros(0, np.float64), np.zeros(0, np.int32)
    span = int(post_dl.max()) + 1
    pair = post_tf * span + post_dl
    uniq, inverse = np.unique(pair, return_inverse=True)
    tf = uniq // span
    dl = uniq % span
    k1, b, avgdl = stats.k1, stats.b, stats.avgdl
    if stats.variant == "okapi":
        g = tf * (k1 + 1) / (tf + k1 * (1 - b + b * dl / avgdl))
    else:
        g = (tf * (k1 + 1)) / (tf + k1 * (1 - b + b * (dl / avgdl)))
    return g.astype(np.float64), inverse.astype(np.int32)


@dataclass
class BM25Arrays:
    stats: BM25Stats
    n_docs_local: int
    term_ptr: np.ndarray
    post_doc: np.ndarray
    post_gcode: np.ndarray
    gtab: np.ndarray
    doc_len: np.ndarray = field(default=None)


def build_arrays(doc_ptr: np.ndarray, doc_tok: np.ndarray, n_terms: int, variant: str = "okapi",
                 k1: float = K1_DEFAULT, b: float = B_DEFAULT, epsilon: float = EPS_DEFAULT,
                 stats: Optional[BM25Stats] = None) -> BM25Arrays:
    """Token-id corpus (CSR over docs) -> arrays for ``engine.BM25Index``.

    ``stats`` carries corpus-GLOBAL statistics when this is one shard of a larger
    corpus; otherwise they are computed from this corpus alone."""
    doc_ptr = np.asarray(doc_ptr, dtype=np.int64)
    doc_tok = np.asarray(doc_tok, dtype=np.int64)
    n = len(doc_ptr) - 1
    doc_len = np.diff(doc_ptr)
    term_ptr, post_doc, post_tf = term_doc_counts(doc_ptr, doc_tok, n_terms)
    if stats is None:
        df = np.diff(term_ptr)
        stats = compute_stats(variant, n, int(doc_len.sum()), df, first_appearance_order(doc_tok, n_terms).tolist(),
                              k1, b, epsilon)
    gtab, gcode = factor_table(stats, post_tf, doc_len[post_doc])
    if len(gtab) > (1 << 18):
        raise ValueError("more than 2^18 distinct (tf, doc_len) pairs; documents this long are outside the index format")
    return BM25Arrays(stats, n, term_ptr, post_doc, gcode, gtab, doc_len)


class Vocabulary:
    """str -> term id in first-appearance order (the reference keys everything by str)."""

    def __init__(self):
        self.ids: Dict[str, int] = {}

    def encode_docs(self, tokenized: List[List[str]]):
        ptr = np.zeros(len(tokenized) + 1, dtype=np.int64)
        toks: List[int] = []
        ids = self.ids
        for i, doc in enumerate(tokenized):
            for w in doc:
                t = ids.get(w)
                if t is None:
                    t = len(ids)
                    ids[w] = t
                toks.append(t)
            ptr[i + 1] = len(toks)
        return ptr, np.asarray(toks, dtype=np.int64)

    def encode_query(self, tokens: Sequence[str]) -> List[int]:
        get = self.ids.get
        return [get(w, -1) for w in tokens]

    def __len__(self):
        return len(self.ids)
