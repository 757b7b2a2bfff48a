"""Oracle: dense flat inner-product top-k (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED for the faiss part: faiss is an external wheel
(``faiss-gpu-cu12>=1.7.4``, reference ``requirements.txt:9-10``) that is not
installed here.  The contract restated is the documented one of
``faiss.IndexFlatIP.search`` as called at ``vector_store/vector_index.py:223``:
exact fp32 inner products, k largest per query in descending order, int64 ids,
``-1`` ids (score = lowest float) beyond ``ntotal``.  faiss leaves the order of
equal scores unspecified; the oracle fixes "descending score, ascending id".
"""
from __future__ import annotations

import numpy as np

LOWEST_F32 = np.float32(-3.4028234663852886e38)


def preprocess_vectors(vectors: np.ndarray, similarity_metric: str = "cosine") -> np.ndarray:
    """``VectorIndex._preprocess_vectors`` (vector_store/vector_index.py:265-282).

    fp32 cast (:268-269), C-contiguous (:272-273), row L2 normalisation with
    ``norm == 0 -> 1`` when the metric is cosine (:276-280).
    """
    if vectors.dtype != np.float32:
        vectors = vectors.astype(np.float32)
    if not vectors.flags["C_CONTIGUOUS"]:
        vectors = np.ascontiguousarray(vectors)
    if similarity_metric == "cosine":
        norms = np.linalg.norm(vectors, axis=1, keepdims=True)
        norms = np.where(norms == 0, 1, norms)
        vectors = vectors / norms
    return vectors


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (round-to-nearest-even) and up-cast back to fp32.

    This is what the bf16 storage mode of the index holds; the oracle for that
    mode scores the rounded values in fp32.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    # NaN stays NaN (not produced by the generators; kept for completeness)
    out = rounded.astype(np.uint32).view(np.float32)
    return np.where(np.isnan(x), x, out).astype(np.float32)


def split_bf16_planes(x: np.ndarray, planes: int = 3):
    """Exact hi/mid/lo bf16 decomposition of fp32 values (x == hi+mid+lo).

    Mirrors the index's fp32-exact storage (3 bf16 planes, truncation split) so
    tests can reason about what the tensor cores are fed.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = []
    rem = x.copy()
    for _ in range(planes):
        p = (rem.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
        out.append(p.copy())
        rem = (rem - p).astype(np.float32)
    return out


def flat_ip_scores(q: np.ndarray, x: np.ndarray) -> np.ndarray:
    """fp32 ``Q @ X.T`` (IndexFlatIP distance computation)."""
    return np.asarray(q, dtype=np.float32) @ np.asarray(x, dtype=np.float32).T


def topk_desc(scores: np.ndarray, k: int):
    """k largest per row, descending score, ascending id on ties; -1 padding."""
    nq, n = scores.shape
    out_s = np.full((nq, k), LOWEST_F32, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    kk = min(k, n)
    if kk == 0:
        return out_s, out_i
    ids = np.arange(n, dtype=np.int64)
    for r in range(nq):
        row = scores[r]
        if kk < n:
            part = np.argpartition(-row, kk - 1)[:kk]
            kth = row[part].min()
            # take every id tied with the k-th score so the tie rule is exact
            cand = np.nonzero(row >= kth)[0]
        else:
            cand = ids
        order = np.lexsort((cand, -row[cand].astype(np.float64)))[:kk]
        sel = cand[order]
        out_s[r, :kk] = row[sel]
        out_i[r, :kk] = sel
    return out_s, out_i


def flat_ip_topk(q: np.ndarray, x: np.ndarray, k: int, block: int = 65536):
    """IndexFlatIP.search restated: blocked fp32 GEMM + per-row top-k.

    Blocked over database rows like faiss's CPU path (which blocks 4096 queries x
    1024 rows, vector_index.py:223 -> faiss); the block size only bounds memory,
    the result does not depend on it except through fp32 summation order inside
    BLAS, which faiss also leaves implementation-defined.
    """
    q = np.asarray(q, dtype=np.float32)
    x = np.asarray(x, dtype=np.float32)
    nq, n = q.shape[0], x.shape[0]
    best_s = np.full((nq, 0), LOWEST_F32, dtype=np.float32)
    best_i = np.full((nq, 0), -1, dtype=np.int64)
    for start in range(0, n, block):
        sc = q @ x[start:start + block].T
        s, i = topk_desc(sc, min(k, sc.shape[1]))
        i = i + start
        cat_s = np.concatenate([best_s, s], axis=1)
        cat_i = np.concatenate([best_i, i], axis=1)
        kk = min(k, cat_s.shape[1])
        new_s = np.empty((nq, kk), np.float32)
        new_i = np.empty((nq, kk), np.int64)
        for r in range(nq):
            order = np.lexsort((cat_i[r], -cat_s[r].astype(np.float64)))[:kk]
            new_s[r] = cat_s[r, order]
            new_i[r] = cat_i[r, order]
        best_s, best_i = new_s, new_i
    out_s = np.full((nq, k), LOWEST_F32, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    out_s[:, :best_s.shape[1]] = best_s
    out_i[:, :best_i.shape[1]] = best_i
    return out_s, out_i


def vector_index_search(query_vectors: np.ndarray, corpus_vectors: np.ndarray, top_k: int = 10,
                        similarity_metric: str = "cosine", total_vectors=None):
    """``VectorIndex.search`` result shaping (vector_store/vector_index.py:206-263).

    ``corpus_vectors`` are the rows as passed to ``add_vectors`` (pre-processing is
    applied here as at :184 and :220).  One query returns a flat list (:256-257).
    """
    if corpus_vectors is None or len(corpus_vectors) == 0:
        return []
    xq = preprocess_vectors(np.asarray(query_vectors), similarity_metric)
    xb = preprocess_vectors(np.asarray(corpus_vectors), similarity_metric)
    scores, indices = flat_ip_topk(xq, xb, top_k)
    results = []
    for qi in range(len(xq)):
        qres = []
        for rank in range(top_k):
            idx = indices[qi][rank]
            score = scores[qi][rank]
            if idx == -1:
                continue
            res = {"index": int(idx), "score": float(score), "rank": rank}
            if similarity_metric == "cosine":
                res["similarity"] = float(score)
            else:
                res["similarity"] = 1.0 / (1.0 + float(score))
            qres.append(res)
        results.append(qres)
    if len(results) == 1:
        return results[0]
    return results


def compare_topk(ids_a, sc_a, ids_b, sc_b, tol: float = 1e-5):
    """Compare two top-k results modulo ties within ``tol`` (north-star rule).

    Returns (ok, n_exact_id_match, n_total, worst_score_gap).  A position may hold
    different ids only if both scores at that position, and the other list's score
    for the swapped id (when present), are within ``tol``.
    """
    ids_a = np.asarray(ids_a); ids_b = np.asarray(ids_b)
    sc_a = np.asarray(sc_a, dtype=np.float64); sc_b = np.asarray(sc_b, dtype=np.float64)
    assert ids_a.shape == ids_b.shape
    same = ids_a == ids_b
    n_match = int(same.sum())
    worst = 0.0
    ok = True
    valid = (ids_a >= 0) | (ids_b >= 0)
    gap = np.abs(sc_a - sc_b)
    gap = np.where(valid, gap, 0.0)
    if gap.size:
        worst = float(gap.max())
    if np.any(gap > tol):
        ok = False
    # where ids differ, the id must be a tie: its score within tol of the k-th score
    # of the other list or present elsewhere in the other list's row
    for r in np.nonzero(~same.all(axis=1))[0]:
        sa = set(ids_a[r].tolist()); sb = set(ids_b[r].tolist())
        only_a = sa - sb
        only_b = sb - sa
        kth_a = sc_a[r][ids_a[r] >= 0].min() if np.any(ids_a[r] >= 0) else 0.0
        kth_b = sc_b[r][ids_b[r] >= 0].min() if np.any(ids_b[r] >= 0) else 0.0
        for i in only_a:
            s = sc_a[r][ids_a[r] == i][0]
            if abs(s - kth_b) > tol:
                ok = False
        for i in only_b:
            s = sc_b[r][ids_b[r] == i][0]
            if abs(s - kth_a) > tol:
                ok = False
    return ok, n_match, int(ids_a.size), worst


# --------------------------------------------------------------------------------------
# IVF (PARITY UNPINNED: faiss.IndexIVFFlat is external and absent, see the module header)
# --------------------------------------------------------------------------------------
def ivf_assign(x: np.ndarray, centroids: np.ndarray) -> np.ndarray:
    """Coarse quantiser of ``faiss.IndexIVFFlat(IndexFlatIP(d), d, nlist, METRIC_INNER_PRODUCT)``
    (created at vector_store/vector_index.py:85-91): a row goes to the list whose centroid has the
    largest inner product with it (lowest list id on exact ties)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = np.ascontiguousarray(centroids, dtype=np.float32)
    out = np.empty(len(x), dtype=np.int32)
    for lo in range(0, len(x), 65536):
        out[lo:lo + 65536] = np.argmax(x[lo:lo + 65536] @ c.T, axis=1)
    return out


def ivf_search(q: np.ndarray, x: np.ndarray, centroids: np.ndarray, assign: np.ndarray, k: int, nprobe: int):
    """``index.search(q, k)`` of an IndexIVFFlat with ``index.nprobe = nprobe`` (vector_index.py:91,:223):
    the ``nprobe`` lists of largest centroid inner product are scanned exhaustively, exact fp32 inner
    products, k best in descending order (ascending id on ties), ``-1`` / lowest float when the probed
    lists hold fewer than k rows.  Returns (scores [nq,k], ids [nq,k], probed lists [nq,nprobe])."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = np.ascontiguousarray(centroids, dtype=np.float32)
    nprobe = min(nprobe, len(c))
    _, probes = topk_desc(q @ c.T, nprobe)
    order = np.argsort(assign, kind="stable")
    starts = np.searchsorted(assign[order], np.arange(len(c) + 1))
    scores = np.full((len(q), k), LOWEST_F32, dtype=np.float32)
    ids = np.full((len(q), k), -1, dtype=np.int64)
    for r in range(len(q)):
        rows = np.concatenate([order[starts[l]:starts[l + 1]] for l in probes[r]]) if nprobe else np.zeros(0, np.int64)
        if len(rows) == 0:
            continue
        rows = np.sort(rows)
        s = x[rows] @ q[r]
        ss, ii = topk_desc(s[None, :], min(k, len(rows)))
        n = ss.shape[1]
        scores[r, :n] = ss[0]
        ids[r, :n] = rows[ii[0]]
    return scores, ids, probes
