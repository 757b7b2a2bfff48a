/*
 * anorag_b200.h -- C ABI of the B200-native candidate-scoring library
 * (libanorag_b200.so, built from ano-rag_b200/csrc for sm_100a).
 *
 * The reference (Kevinwu901113/ano-rag) is pure Python; its "FFI" for this path
 * is the set of calls it makes into the external faiss / rank_bm25 packages and
 * its own pure-Python fusion.  Each entry point below replaces one of those call
 * sites (reference file:line given per function).  The Python classes in
 * ano-rag_b200/ (VectorIndex, VectorRetriever, bm25_search, HybridSearcher) bind
 * these symbols with ctypes -- see INTEGRATION.md for the stub.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - every function returns 0 on success, a negative ARB_E* code otherwise;
 *     arb_last_error() returns a thread-local message for the last failure.
 *   - `*_is_device` says whether a buffer is device memory (1) or host memory (0);
 *     host buffers are staged by the library with cudaMemcpyAsync on `stream`
 *     (pinned host memory keeps that asynchronous).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Work is
 *     enqueued on it; functions that write HOST outputs synchronise the stream
 *     before returning, functions with device outputs do not.
 *   - handles are opaque.  Searches on a handle are re-entrant (per-call,
 *     stream-ordered scratch; the index is read-only during search); create /
 *     add / reset / destroy are single-writer.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with ARB_ECUDA.
 */
#ifndef ANORAG_B200_H_
#define ANORAG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARB_ABI_VERSION 1

#if defined(__GNUC__)
#define ARB_API __attribute__((visibility("default")))
#else
#define ARB_API
#endif

#define ARB_OK        0
#define ARB_EINVAL   -1   /* bad argument                         */
#define ARB_ECUDA    -2   /* CUDA runtime / driver error          */
#define ARB_ENOMEM   -3   /* allocation failure                   */
#define ARB_ELIMIT   -4   /* argument outside a documented limit  */

/* element types of caller buffers */
#define ARB_F32   0
#define ARB_BF16  1
#define ARB_F64   2

/* dense storage modes */
#define ARB_STORE_BF16       0  /* one bf16 plane per row: bf16(normalised row)            */
#define ARB_STORE_FP32_SPLIT 1  /* three bf16 planes hi/mid/lo, hi+mid+lo == fp32 row      */

/* fusion methods (retrieval/hybrid_search.py:22) */
#define ARB_FUSE_LINEAR 0
#define ARB_FUSE_RRF    1

/* BM25 variants */
#define ARB_BM25_OKAPI  0  /* rank_bm25.BM25Okapi arithmetic  (b*dl/avgdl)                 */
#define ARB_BM25_SIMPLE 1  /* utils/bm25_search.py:16-63      (b*(dl/avgdl))               */

#define ARB_MAX_TOPK        256   /* dense and BM25 top-k per call                          */
#define ARB_MAX_QUERY_TERMS 64    /* BM25 query tokens per query (duplicates count)         */

typedef struct arb_index arb_index;   /* dense flat inner-product index       */
typedef struct arb_bm25  arb_bm25;    /* BM25 block-partitioned inverted index */

ARB_API int         arb_abi_version(void);
ARB_API const char* arb_last_error(void);
/* number of CUDA devices visible, or a negative ARB_E* code */
ARB_API int         arb_device_count(void);
/* kernels launched by this library on the calling thread since the last reset
 * (bench.py reports it as gpu_launches) */
ARB_API int64_t     arb_launch_count(int reset);
/* Per-thread kernel timing for the roofline report: while enabled, the dominant
 * kernels are bracketed by CUDA events on their launch stream.  collect() waits for
 * them and returns summed milliseconds / launch counts per kind, then clears:
 * [0] dense GEMM+select  [1] BM25 top-k (whole arb_bm25_search)  [2] fusion  [3] dense list merge
 * [4] BM25 sparse pass  [5] BM25 plane pass  [6] BM25 final merge + fallback  [7] IVF coarse probe. */
#define ARB_PROFILE_KINDS 8
ARB_API int         arb_profile(int enable);
ARB_API int         arb_profile_collect(double out_ms[ARB_PROFILE_KINDS], int64_t out_count[ARB_PROFILE_KINDS]);

/* ------------------------------------------------------------------------- *
 * Dense: exact inner-product top-k (replaces faiss.IndexFlatIP behind
 * VectorIndex -- vector_store/vector_index.py:79 create, :187-196 add,
 * :223 search, :265-282 pre-processing).
 * ------------------------------------------------------------------------- */

/* vector_index.py:61-124 (create_index 'Flat' + cosine/IP).  `normalize` != 0
 * applies the row L2 normalisation of _preprocess_vectors (:276-280) to added
 * rows and to queries on the device. */
ARB_API int arb_index_create(arb_index** out, int dim, int storage, int normalize);
ARB_API int arb_index_destroy(arb_index* index);
/* vector_index.py:171-204 add_vectors: append `n` rows of `dim` elements
 * (row-major, `dtype` ARB_F32 or ARB_BF16).  Row order is the id (:194-196). */
ARB_API int arb_index_add(arb_index* index, const void* rows, int dtype, int64_t n,
                  int rows_is_device, void* stream);
/* Rows of the following add calls are (on != 0) / are not L2-normalised on the device; queries keep the
 * setting of arb_index_create.  load_index (vector_index.py:323-364) uses it to take rows that a saved
 * index already holds normalised bit for bit, as faiss.read_index does. */
ARB_API int arb_index_set_row_normalization(arb_index* index, int on);
ARB_API int arb_index_reset(arb_index* index);                    /* :409-421 reset_index */
ARB_API int64_t arb_index_ntotal(const arb_index* index);         /* faiss ntotal (:381)   */
ARB_API int arb_index_dim(const arb_index* index);
/* stored rows back as fp32 (sum of planes) -- used by save_index (:284-321) */
ARB_API int arb_index_reconstruct(const arb_index* index, int64_t row0, int64_t n, float* out,
                          int out_is_device, void* stream);

/* vector_index.py:223 `scores, indices = self.index.search(q, k)`.
 *   queries       [nq, dim] row-major, ARB_F32 or ARB_BF16
 *   query_planes  1 = queries rounded to bf16 (one tensor-core pass per plane pair)
 *                 3 = queries split hi/mid/lo (fp32-exact queries)
 *   out_scores    [nq, k] fp32, descending; slots beyond ntotal hold -3.4028235e38
 *   out_ids       [nq, k] int64 = id_base + row; -1 beyond ntotal (faiss contract)
 * Ties: equal scores are ordered by ascending id.  k <= ARB_MAX_TOPK. */
ARB_API int arb_index_search(const arb_index* index, const void* queries, int dtype, int64_t nq,
                     int queries_is_device, int k, int query_planes, int64_t id_base,
                     float* out_scores, int64_t* out_ids, int out_is_device, void* stream);

/* Per-candidate similarities (replaces the loop of QueryProcessor._calculate_vector_similarities,
 * query/query_processor.py:3543-3589): out[q, c] = inner product of query q (pre-processed like a
 * search query: L2-normalised when the index normalises) with stored row cand_rows[q, c]; with
 * clamp_nonneg != 0 the result is max(0, .) (:3567); rows < 0 or >= ntotal score 0. */
ARB_API int arb_index_score_rows(const arb_index* index, const void* queries, int dtype, int64_t nq,
                         int queries_is_device, const int64_t* cand_rows, int nc, int cand_is_device,
                         int query_planes, int clamp_nonneg, float* out, int out_is_device, void* stream);

/* ------------------------------------------------------------------------- *
 * IVF: faiss.IndexIVFFlat(IndexFlatIP(d), d, nlist, METRIC_INNER_PRODUCT) behind VectorIndex
 * (vector_index.py:85-91 create, :126-169 train_index, :187-193 add_with_ids, :223 search with
 * index.nprobe).  An index becomes an IVF index once it has centroids; rows added before or after
 * are quantised to the centroid of largest inner product (IndexFlatIP quantiser), ids stay the add
 * order.  The flat search (arb_index_search) keeps working on the same handle.
 * ------------------------------------------------------------------------- */
/* install `nlist` centroids [nlist, dim] fp32 as given (e.g. topic-cluster means, doc/clustering.py:194) */
ARB_API int arb_index_ivf_set_centroids(arb_index* index, int nlist, const float* centroids,
                                int centroids_is_device, void* stream);
/* index.train(x) (:158): k-means on the device -- seeds = evenly spaced training rows, `niter` Lloyd
 * iterations (faiss default 25), `spherical` != 0 re-normalises the centroids (faiss does for IP). */
ARB_API int arb_index_ivf_train(arb_index* index, int nlist, const void* rows, int dtype, int64_t n,
                        int rows_is_device, int niter, int spherical, void* stream);
ARB_API int arb_index_ivf_nlist(const arb_index* index);                /* 0 = not trained */
/* centroids [nlist, dim] fp32 and/or the list of every stored row int32 [ntotal] (either may be NULL); host */
ARB_API int arb_index_ivf_export(const arb_index* index, float* centroids_out, int32_t* assign_out, void* stream);
/* search the `nprobe` lists of largest centroid inner product; same outputs as arb_index_search
 * (rows outside the probed lists are not seen: slots stay -1 / -3.4028235e38 when fewer than k). */
ARB_API int arb_index_ivf_search(arb_index* index, const void* queries, int dtype, int64_t nq,
                         int queries_is_device, int k, int nprobe, int query_planes, int64_t id_base,
                         float* out_scores, int64_t* out_ids, int out_is_device, void* stream);

/* ------------------------------------------------------------------------- *
 * BM25 (replaces corpus.get_scores at utils/bm25_search.py:309 /
 * SimpleBM25.get_scores :43-63, and the max-normalisation :329-333).
 *
 * The host builds the term dictionary (strings never reach the device) and
 * passes a term-major CSR: postings of term t are
 * [term_ptr[t], term_ptr[t+1]) sorted by ascending doc id; post_gcode indexes
 * `gtab`, the table of distinct per-(tf, doc_len) factors
 *     g = tf*(k1+1) / (tf + k1*(1 - b + b*dl/avgdl))        (fp64, reference order)
 * so that a posting's contribution is exactly idf[t] * g, the two roundings the
 * reference performs.  All host pointers.  n_docs <= 2^31-1, n_g <= 2^18.
 * ------------------------------------------------------------------------- */
ARB_API int arb_bm25_create(arb_bm25** out, int64_t n_docs, int64_t n_terms,
                    const int64_t* term_ptr, const int32_t* post_doc, const int32_t* post_gcode,
                    const double* idf, const double* gtab, int64_t n_g, void* stream);
ARB_API int arb_bm25_destroy(arb_bm25* bm);
ARB_API int64_t arb_bm25_ndocs(const arb_bm25* bm);
ARB_API int64_t arb_bm25_npostings(const arb_bm25* bm);
/* sum over the given query terms (with multiplicity) of df(t): the algorithmic
 * posting count the roofline figure is computed from.  Host q_terms. */
ARB_API int64_t arb_bm25_query_postings(const arb_bm25* bm, const int32_t* q_terms, int64_t nq, int max_qt);

/* Batched top-k.  q_terms [nq, max_qt] int32 term ids in QUERY TOKEN ORDER,
 * duplicates kept, -1 = padding / unknown token.  Per query, per doc:
 * score = (((0 + c_1) + c_2) + ...) in fp64, c_j = idf[t_j]*g(posting), exactly
 * the order of rank_bm25 get_scores.  Output: descending score, ascending doc id
 * on ties, docs with score <= 0 excluded, ids = id_base + doc, -1 / 0.0 padding. */
ARB_API int arb_bm25_search(const arb_bm25* bm, const int32_t* q_terms, int64_t nq, int max_qt,
                    int q_is_device, int k, int64_t id_base, double* out_scores, int64_t* out_ids,
                    int out_is_device, void* stream);
/* Full score vectors (bm25_scores, utils/bm25_search.py:286-340): out [nq, n_out]
 * fp64; docs >= n_out are truncated, slots >= n_docs padded with 0.0 (:320-327);
 * bit 0 of `normalise` divides by the per-query max when it is > 0 (:329-333); bit 1 CONTINUES the sums already in
 * `out` (device-resident) instead of starting from zero: a query longer than ARB_MAX_QUERY_TERMS is scored in
 * consecutive chunks, which keeps the reference's token-by-token summation order bit for bit. */
ARB_API int arb_bm25_scores(const arb_bm25* bm, const int32_t* q_terms, int64_t nq, int max_qt,
                    int q_is_device, int normalise, int64_t n_out, double* out,
                    int out_is_device, void* stream);

/* Diagnostics of the batched pipeline, collected only while the environment variable ARB_BM25_STATS is
 * set (each search then synchronises): [0] queries scored, [1] queries re-scored by the per-query
 * fallback, [2] screening-pass emissions in total, [3] most emissions for one query, [4] (query, tile) pairs the
 * screening pass recomputed in fp64, [5] (record, doc) pairs the member pass set aside for the exact kernel, [6] those
 * of them that did match another sparse token, [7] reserved.  Per calling thread. */
ARB_API int arb_bm25_batch_stats(int64_t out[8], int reset);

/* FieldWeightedBM25.get_scores combination (utils/bm25_search.py:117-145):
 * out[i] = ((0.0 + w[0]*s[0][i]) + w[1]*s[1][i]) + ... in fp64, fields in the given order; then, if
 * `nq` > 0, every row of length n/nq is divided by its max when that is > 0 (:219-223).  Device buffers. */
ARB_API int arb_combine_fields(int n_fields, const double* const* field_scores, const double* weights,
                               int64_t n, int64_t nq, double* out, void* stream);

/* ------------------------------------------------------------------------- *
 * Fusion (replaces HybridSearcher.fuse, retrieval/hybrid_search.py:34-103).
 * Each source s in {dense, bm25, graph, path} is a [nq, k_s] list of
 * (id int64, score) with id < 0 = empty slot; ids are unique inside a list.
 * `score_dtype[s]` is ARB_F32 or ARB_F64.  Arithmetic is fp64 in the reference's
 * operation order (no fused multiply-add).  Output per query: up to `pool`
 * entries, descending final score, exact ties in first-appearance order of the
 * dense -> bm25 -> graph -> path traversal:
 *   out_ids   [nq, pool] int64 (-1 padding)
 *   out_final [nq, pool] fp64
 *   out_src   [nq, pool, 4] fp64 raw per-source score, NaN where absent
 *   out_count [nq] int32
 * All buffers device memory (is_device=1) or all host (0).  k_s <= 256 each. */
ARB_API int arb_fuse(int method, const double weights[4], double rrf_k, int pool, int64_t nq,
             const int64_t* const ids[4], const void* const scores[4], const int score_dtype[4],
             const int k_src[4], int64_t* out_ids, double* out_final, double* out_src,
             int32_t* out_count, int is_device, void* stream);

/* Candidate re-scoring of the live pipeline (replaces the fusion arithmetic and the sort of
 * QueryProcessor._hybrid_search, query/query_processor.py:3709-3760).  Per query q, candidates
 * i < n_cand[q] (NULL: all nc): `vec`/`bm25` [nq, nc] fp64 (missing scores = 0.0, :3712-3713);
 * guard flags computed by the host from the note text (:3717-3733): must_miss (no must-have term in
 * the content), n_ent / n_pred (matching boost entities / predicates, with multiplicity).
 *   ARB_FUSE_LINEAR: b *= 0.1 if must_miss; v *= 1.2 per entity; b *= 1.3 per predicate; vw*v + bw*b
 *   ARB_FUSE_RRF:    vw/(rrf_k + rank_v) + bw/(rrf_k + rank_b), 0-based stable descending ranks;
 *                    *= 0.1 if must_miss
 * out_score [nq, nc] in candidate order; out_order [nq, nc] = candidate indices by descending score
 * (stable), -1 beyond n_cand.  All pointers are device pointers; nc <= 4096. */
ARB_API int arb_rescore_candidates(int method, int64_t nq, int nc, const double* vec, const double* bm25,
                           const int32_t* n_cand, const uint8_t* must_miss, const uint8_t* n_ent,
                           const uint8_t* n_pred, double vector_weight, double bm25_weight, double rrf_k,
                           double* out_score, int32_t* out_order, void* stream);

/* QueryProcessor._enhanced_hybrid_search_v2 (reference query/query_processor.py:1089-1143, without the optional
 * ListT5 stage), batched: per candidate
 *   f = dense_weight*dense + sparse_weight*sparse          (1.0 / 0.6 in the reference, :1109)
 *   f *= section_mult; f *= lexical_mult                   (_apply_section_filtering :1235, _apply_lexical_fallback :1259)
 *   if (f < noise_threshold && !satisfies_must) f = 0      (:1122-1123)
 *   f *= entity_boost; f *= predicate_boost                (:1126-1132)
 * candidates with f > 0 are kept and sorted by f, descending and stable (:1139-1142).  The multipliers come
 * from note text and are computed by the host; a NULL multiplier array (or 1.0) skips the step, a NULL
 * satisfies_must means "every candidate satisfies".  All arrays [nq, nc] device memory, nc <= 4096.
 * out_score [nq, nc] in candidate order; out_order [nq, nc] = kept candidate indices in result order, -1
 * beyond out_kept[q] (out_kept may be NULL). */
ARB_API int arb_rescore_v2(int64_t nq, int nc, const double* dense, const double* sparse, const int32_t* n_cand,
                   const double* section_mult, const double* lexical_mult, const uint8_t* satisfies_must,
                   const double* entity_boost, const double* predicate_boost, double dense_weight,
                   double sparse_weight, double noise_threshold, double* out_score, int32_t* out_order,
                   int32_t* out_kept, void* stream);

/* QueryProcessor._rerank_khop_candidates (reference query/query_processor.py:1549-1635) with _rrf_score
 * (:854-861), batched: per candidate
 *   final = dense_weight*dense + bm25_weight*bm25 + focused_w*focused + cov_weight*cov + cons_weight*cons
 *           + rrf_lambda * (1/(rrf_k + rank_dense) + 1/(rrf_k + rank_bm25) [+ 1/(rrf_k + rank_prf)])
 *   final *= hop_mult                                       (hop_decay ** (hop_no - 1), 1.0 when not applied)
 * ranks are 1-based, stable, descending; rank_prf is taken inside the subset is_prf != 0, ordered by prf_sim.
 * `group` (NULL = every candidate its own) carries the note id: candidates of one group share one RRF
 * entry holding the largest rank of the group, as the reference's dict keyed by note_id does.  focused_w
 * (per candidate, from hop_no) and hop_mult are computed by the host.  out_order: all n_cand[q] indices by
 * descending final (stable), -1 beyond. */
ARB_API int arb_rescore_khop(int64_t nq, int nc, const double* dense, const double* bm25, const double* focused,
                     const double* cov, const double* cons, const double* prf_sim, const uint8_t* is_prf,
                     const int32_t* group, const double* focused_w, const double* hop_mult,
                     const int32_t* n_cand, double dense_weight, double bm25_weight, double cov_weight,
                     double cons_weight, double rrf_lambda, int rrf_k, double* out_score, int32_t* out_order,
                     void* stream);

/* ------------------------------------------------------------------------- *
 * TF-IDF cosine "BM25 fallback" of VectorRetriever (reference vector_store/retriever.py:924-1002):
 * cosine_similarity(query_vector, tfidf_matrix) (:977) + argsort top-k (:980) + `similarity > 0` (:985).
 * The host keeps sklearn's TfidfVectorizer; the library holds the L2-normalised document matrix term-major
 * (CSC of the doc x term matrix: for term t the postings term_ptr[t] .. term_ptr[t+1], doc ids strictly
 * ascending) and accumulates exactly like scipy's sparse product: per doc, query_weight * doc_weight over the
 * query's terms in ascending term order, every product and sum rounded separately.
 * ------------------------------------------------------------------------- */
typedef struct arb_tfidf arb_tfidf;
ARB_API int arb_tfidf_create(arb_tfidf** out, int64_t n_docs, int64_t n_terms, const int64_t* term_ptr,
                     const int32_t* post_doc, const double* post_val);          /* host arrays */
ARB_API int arb_tfidf_destroy(arb_tfidf* index);
/* nq sparse, L2-normalised queries in CSR form (q_term ascending inside a query); host arrays in and out.
 * out_scores / out_ids [nq, k]: descending similarity; among EQUAL similarities the larger doc id first (the reference's
 * np.argsort is an unstable sort, so it leaves that order unspecified); only similarities > 0; id -1 / score 0 beyond. */
ARB_API int arb_tfidf_search(const arb_tfidf* index, int64_t nq, const int64_t* q_ptr, const int32_t* q_term,
                     const double* q_val, int k, double* out_scores, int64_t* out_ids, void* stream);

/* ------------------------------------------------------------------------- *
 * Multi-GPU merge (new design): `parts` per-shard top-k lists,
 * laid out [parts, nq, k] as an NCCL all-gather leaves them, merged to the global
 * top-k (descending score, ascending id; id < 0 = empty).  Device buffers.
 * ------------------------------------------------------------------------- */
ARB_API int arb_merge_topk(int score_dtype, const void* scores, const int64_t* ids, int parts,
                   int64_t nq, int k, int positive_only, void* out_scores, int64_t* out_ids,
                   void* stream);

/* The same merge without the NCCL exchange (new design): part p's lists start at part_scores[p] /
 * part_ids[p] -- one [rows, k] array per GPU, peer-mapped (CUDA IPC / symmetric memory), so the kernel
 * loads the other shards' candidates straight over NVLink -- and query q of this call is row row0 + q
 * of every part.  part_scores / part_ids are HOST arrays of `parts` (<= 16) device pointers.  The caller
 * orders the peers' writes before this launch (a cross-GPU barrier on `stream`). */
ARB_API int arb_merge_topk_peers(int score_dtype, const void* const* part_scores, const int64_t* const* part_ids,
                         int parts, int64_t row0, int64_t nq, int k, void* out_scores, int64_t* out_ids,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ANORAG_B200_H_ */
