// Score fusion and top-k merging (sm_100a).
//
// fuse_kernel reproduces HybridSearcher.fuse (reference
// retrieval/hybrid_search.py:34-103) for a batch of queries: per-source
// max-normalisation + weighted sum ("linear", :83-100) or weighted reciprocal-rank
// fusion ("rrf", :64-82), stable descending sort (:102) and the candidate_pool cut
// (:103).  All arithmetic is fp64 with explicit round-to-nearest intrinsics so the
// compiler cannot contract a*b+c into an FMA: Python evaluates each product and
// sum separately.
//
// merge_topk_kernel merges per-shard / per-split top-k lists (new design, the
// reference has no sharding): one CTA per query, bitonic sort in shared memory.
#include "common.cuh"

#include <algorithm>
#include <mutex>

namespace arb {

constexpr int FUSE_THREADS = 256;
constexpr int FUSE_MAXK = 256;                 // entries per source (largest instantiation)
constexpr int FUSE_MAXU = 4 * FUSE_MAXK;       // union size bound (also sort size, power of two)

struct FuseParams {
  int method;
  double w[4];
  double rrf_k;
  int pool;
  int64_t nq;
  const int64_t* ids[4];
  const void* scores[4];
  int dtype[4];
  int k[4];
  int64_t* out_ids;
  double* out_final;
  double* out_src;
  int32_t* out_count;
};

// Shared memory of one query, sized by the per-source capacity MAXK (128 or 256): 41 KB / 82 KB, i.e. five / two
// CTAs per SM -- the kernel is latency-bound (barriers of a small sort), so residency is what it scales with.
template <int MAXK>
struct FuseSmem {
  static constexpr int HASH = 8 * MAXK;        // hash slots: load factor <= 0.5 for 4 full sources
  static constexpr int MAXU = 4 * MAXK;        // union size bound (also sort size, power of two)
  int64_t ids[4][MAXK];
  double sc[4][MAXK];
  short rank[4][MAXK];               // 1-based stable descending rank inside the source (rrf)
  unsigned long long hkey[HASH];
  short hpos[HASH][4];               // position of the id in each source, -1 = absent
  unsigned long long skey[MAXU];
  uint32_t sseq[MAXU];
  uint32_t sslot[MAXU];
  double hfin[HASH];                 // final score of the id in hash slot h
  double maxv[4];
  int cnt[4];
  int n_union;
  double red[FUSE_THREADS / 32];
};

constexpr unsigned long long HASH_EMPTY = 0xFFFFFFFFFFFFFFFFull;

template <int HASH>
__device__ __forceinline__ uint32_t hash_id(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return static_cast<uint32_t>(x) & (HASH - 1);
}

template <int MAXK>
__global__ void __launch_bounds__(FUSE_THREADS)
fuse_kernel(const FuseParams p) {
  extern __shared__ __align__(16) uint8_t fsm_raw[];
  using Smem = FuseSmem<MAXK>;
  constexpr int FUSE_HASH = Smem::HASH;
  constexpr int FUSE_MAXU = Smem::MAXU;
  Smem& s = *reinterpret_cast<Smem*>(fsm_raw);
  const int tid = threadIdx.x;
  const int64_t q = blockIdx.x;

  // ---- load the four lists
  for (int src = 0; src < 4; ++src) {
    const int ks = p.k[src];
    for (int j = tid; j < ks; j += FUSE_THREADS) {
      const int64_t id = p.ids[src] ? p.ids[src][q * ks + j] : -1;
      double v = 0.0;
      if (id >= 0) {
        v = p.dtype[src] == ARB_F32 ? static_cast<double>(static_cast<const float*>(p.scores[src])[q * ks + j])
                                    : static_cast<const double*>(p.scores[src])[q * ks + j];
      }
      s.ids[src][j] = id;
      s.sc[src][j] = v;
    }
  }
  for (int i = tid; i < FUSE_HASH; i += FUSE_THREADS) {
    s.hkey[i] = HASH_EMPTY;
    s.hpos[i][0] = s.hpos[i][1] = s.hpos[i][2] = s.hpos[i][3] = -1;
  }
  if (tid < 4) { s.cnt[tid] = 0; s.maxv[tid] = 0.0; }
  if (tid == 0) s.n_union = 0;
  __syncthreads();

  // ---- per-source valid count and max (hybrid_search.py:29 `max(scores.values())`)
  for (int src = 0; src < 4; ++src) {
    const int ks = p.k[src];
    double m = -INFINITY; int c = 0;
    for (int j = tid; j < ks; j += FUSE_THREADS)
      if (s.ids[src][j] >= 0) { m = fmax(m, s.sc[src][j]); ++c; }
    for (int o = 16; o > 0; o >>= 1) { m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o)); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if ((tid & 31) == 0) { s.red[tid >> 5] = m; atomicAdd(&s.cnt[src], c); }
    __syncthreads();
    if (tid == 0) {
      double mm = -INFINITY;
      for (int w = 0; w < FUSE_THREADS / 32; ++w) mm = fmax(mm, s.red[w]);
      s.maxv[src] = mm;
    }
    __syncthreads();
  }

  // ---- stable descending rank inside each source (sorted(..., reverse=True) keeps
  //      the original order of equal scores, hybrid_search.py:67)
  for (int src = 0; src < 3; ++src) {
    const int ks = p.k[src];
    for (int j = tid; j < ks; j += FUSE_THREADS) {
      if (s.ids[src][j] < 0) { s.rank[src][j] = 0; continue; }
      const double v = s.sc[src][j];
      int r = 1;
      for (int i = 0; i < ks; ++i) {
        if (s.ids[src][i] < 0) continue;
        const double o = s.sc[src][i];
        r += (o > v || (o == v && i < j)) ? 1 : 0;
      }
      s.rank[src][j] = static_cast<short>(r);
    }
  }

  // ---- union of ids via a shared-memory hash table
  for (int src = 0; src < 4; ++src) {
    const int ks = p.k[src];
    for (int j = tid; j < ks; j += FUSE_THREADS) {
      const int64_t id = s.ids[src][j];
      if (id < 0) continue;
      const unsigned long long key = static_cast<unsigned long long>(id);
      uint32_t h = hash_id<FUSE_HASH>(key);
      while (true) {
        const unsigned long long prev = atomicCAS(&s.hkey[h], HASH_EMPTY, key);
        if (prev == HASH_EMPTY || prev == key) {
          // an id that occurs twice in one source keeps its LAST position, like the reference's dict built from the list
          // (retrieval/hybrid_search.py:46-49) -- a 16-bit atomic max through the enclosing 32-bit word, so the result
          // does not depend on which thread gets here first
          unsigned int* w = reinterpret_cast<unsigned int*>(&s.hpos[h][src & ~1]);
          unsigned int seen = *w;
          while (true) {
            const short cur = static_cast<short>((src & 1) ? (seen >> 16) : (seen & 0xFFFFu));
            if (cur >= static_cast<short>(j)) break;
            const unsigned int want = (src & 1) ? ((seen & 0xFFFFu) | (static_cast<unsigned int>(j) << 16))
                                                : ((seen & 0xFFFF0000u) | static_cast<unsigned int>(j));
            const unsigned int got = atomicCAS(w, seen, want);
            if (got == seen) break;
            seen = got;
          }
          break;
        }
        h = (h + 1) & (FUSE_HASH - 1);
      }
    }
  }
  __syncthreads();

  // offsets of each source inside the traversal sequence that fixes tie order
  const int offs[4] = {0, p.k[0], p.k[0] + p.k[1], p.k[0] + p.k[1] + p.k[2]};

  // ---- score every union member
  for (int i = tid; i < FUSE_MAXU; i += FUSE_THREADS) { s.skey[i] = 0ull; s.sseq[i] = 0xFFFFFFFFu; s.sslot[i] = 0xFFFFFFFFu; }
  __syncthreads();
  for (int h = tid; h < FUSE_HASH; h += FUSE_THREADS) {
    if (s.hkey[h] == HASH_EMPTY) continue;
    const short pd = s.hpos[h][0], pb = s.hpos[h][1], pg = s.hpos[h][2], pp = s.hpos[h][3];
    const short pos[4] = {pd, pb, pg, pp};
    double fin;
    uint32_t seq = 0;
    if (p.method == ARB_FUSE_RRF) {
      if (pd < 0 && pb < 0 && pg < 0) continue;                    // path-only ids are not emitted (:64-82)
      double acc = 0.0;                                            // ranks.setdefault(nid, 0.0)
      bool first = true;
      for (int src = 0; src < 3; ++src) {
        if (pos[src] < 0) continue;
        const int r = s.rank[src][pos[src]];
        acc = __dadd_rn(acc, __ddiv_rn(p.w[src], __dadd_rn(p.rrf_k, static_cast<double>(r))));
        if (first) { seq = offs[src] + (r - 1); first = false; }
      }
      const double pv = pp >= 0 ? s.sc[3][pp] : 0.0;
      fin = __dadd_rn(acc, __dmul_rn(p.w[3], pv));
    } else {
      double nv[3];
      for (int src = 0; src < 3; ++src) {
        double v = 0.0;
        if (pos[src] >= 0) {
          const double m = s.maxv[src];
          v = (m == 0.0) ? 0.0 : __ddiv_rn(s.sc[src][pos[src]], m);   // _normalize :26-32
        }
        nv[src] = v;
      }
      const double pv = pp >= 0 ? s.sc[3][pp] : 0.0;
      fin = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(p.w[0], nv[0]), __dmul_rn(p.w[1], nv[1])), __dmul_rn(p.w[2], nv[2])),
                      __dmul_rn(p.w[3], pv));
      for (int src = 0; src < 4; ++src) if (pos[src] >= 0) { seq = offs[src] + pos[src]; break; }
    }
    const int slot = atomicAdd(&s.n_union, 1);
    double cmp = (fin == 0.0) ? 0.0 : fin;                          // -0.0 == 0.0 in Python's sort
    unsigned long long key = f64_key(cmp);
    if (fin != fin) key = 1ull;                                     // NaN sorts last
    s.skey[slot] = key;
    s.sseq[slot] = seq;
    s.sslot[slot] = static_cast<uint32_t>(h);
    s.hfin[h] = fin;
  }
  __syncthreads();

  // ---- stable descending sort: (final desc, traversal sequence asc), over the smallest power of two that
  //      holds the union (the slots beyond it were initialised to "empty", which sorts last)
  int n_sort = 32;
  while (n_sort < s.n_union) n_sort <<= 1;
  for (int size = 2; size <= n_sort; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < n_sort / 2; i += FUSE_THREADS) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool best_first = ((lo & size) == 0);
        const unsigned long long ka = s.skey[lo], kb = s.skey[hi];
        const uint32_t qa = s.sseq[lo], qb = s.sseq[hi];
        const bool a_before_b = (ka > kb) || (ka == kb && qa < qb);
        if (a_before_b != best_first) {
          s.skey[lo] = kb; s.skey[hi] = ka;
          s.sseq[lo] = qb; s.sseq[hi] = qa;
          const uint32_t ta = s.sslot[lo]; s.sslot[lo] = s.sslot[hi]; s.sslot[hi] = ta;
        }
      }
      __syncthreads();
    }
  }

  // ---- emit the pool
  const int n_out = min(s.n_union, p.pool);
  for (int i = tid; i < p.pool; i += FUSE_THREADS) {
    int64_t* oid = p.out_ids + q * p.pool + i;
    double* ofin = p.out_final + q * p.pool + i;
    double* osrc = p.out_src + (q * p.pool + i) * 4;
    if (i < n_out) {
      const uint32_t h = s.sslot[i];
      *oid = static_cast<int64_t>(s.hkey[h]);
      *ofin = s.hfin[h];
      for (int src = 0; src < 4; ++src) {
        const short ps = s.hpos[h][src];
        osrc[src] = ps >= 0 ? s.sc[src][ps] : __longlong_as_double(0x7ff8000000000000ll);
      }
    } else {
      *oid = -1; *ofin = 0.0;
      for (int src = 0; src < 4; ++src) osrc[src] = __longlong_as_double(0x7ff8000000000000ll);
    }
  }
  if (tid == 0) p.out_count[q] = n_out;
}

// ----------------------------------------------------------------------------
// merge: scores/ids [parts, nq, k] -> [nq, k]; descending score, ascending id.
// ----------------------------------------------------------------------------
// With PEERS the lists of part p start at src.scores[p] / src.ids[p] (one [rows, k] array per GPU, read over
// NVLink through peer-mapped symmetric memory) and query q of this launch is row row0 + q of each of them.
constexpr int MERGE_MAX_PEERS = 16;
struct MergePeers {
  const void* scores[MERGE_MAX_PEERS];
  const int64_t* ids[MERGE_MAX_PEERS];
  int64_t row0;
};

template <typename T, bool PEERS>
__global__ void __launch_bounds__(256)
merge_topk_kernel(const T* __restrict__ scores, const int64_t* __restrict__ ids, int parts, int64_t nq, int k,
                  int n_pow2, T* __restrict__ out_scores, int64_t* __restrict__ out_ids, T pad_score,
                  const __grid_constant__ MergePeers src) {
  extern __shared__ __align__(16) uint8_t msm[];
  unsigned long long* key = reinterpret_cast<unsigned long long*>(msm);
  int64_t* id = reinterpret_cast<int64_t*>(msm + static_cast<size_t>(n_pow2) * 8);
  const int tid = threadIdx.x;
  const int64_t q = blockIdx.x;
  const int m = parts * k;
  for (int i = tid; i < n_pow2; i += blockDim.x) {
    unsigned long long kk = 0ull; int64_t ii = 0x7fffffffffffffffll;
    if (i < m) {
      const int part = i / k, j = i % k;
      const int64_t at = PEERS ? (src.row0 + q) * k + j : (static_cast<int64_t>(part) * nq + q) * k + j;
      const int64_t* pi = PEERS ? src.ids[part] : ids;
      const T* ps = PEERS ? static_cast<const T*>(src.scores[part]) : scores;
      const int64_t v = pi[at];
      if (v >= 0) { ii = v; kk = f64_key(static_cast<double>(ps[at])); }
    }
    key[i] = kk; id[i] = ii;
  }
  __syncthreads();
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < n_pow2 / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool best_first = ((lo & size) == 0);
        const unsigned long long ka = key[lo], kb = key[hi];
        const int64_t ia = id[lo], ib = id[hi];
        const bool a_before_b = (ka > kb) || (ka == kb && ia < ib);
        if (a_before_b != best_first) { key[lo] = kb; key[hi] = ka; id[lo] = ib; id[hi] = ia; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += blockDim.x) {
    const bool ok = i < n_pow2 && key[i] != 0ull;
    out_scores[q * k + i] = ok ? static_cast<T>(key_f64(key[i])) : pad_score;
    out_ids[q * k + i] = ok ? id[i] : -1;
  }
}

static int merge_topk_launch(int score_dtype, const void* scores, const int64_t* ids, const MergePeers* peers, int parts,
                             int64_t nq, int k, void* out_scores, int64_t* out_ids, cudaStream_t st) {
  ARB_CHECK(score_dtype == ARB_F32 || score_dtype == ARB_F64, ARB_EINVAL, "merge_topk: score dtype must be f32 or f64");
  ARB_CHECK(parts >= 1 && k >= 1 && nq >= 0, ARB_EINVAL, "merge_topk: bad shape");
  const int64_t m = static_cast<int64_t>(parts) * k;
  ARB_CHECK(m <= 8192, ARB_ELIMIT, "merge_topk: parts * k must be <= 8192");
  if (nq == 0) return ARB_OK;
  int n_pow2 = 2;
  while (n_pow2 < m) n_pow2 <<= 1;
  const size_t smem = static_cast<size_t>(n_pow2) * 16;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    const void* fns[4] = {reinterpret_cast<const void*>(merge_topk_kernel<float, false>),
                          reinterpret_cast<const void*>(merge_topk_kernel<double, false>),
                          reinterpret_cast<const void*>(merge_topk_kernel<float, true>),
                          reinterpret_cast<const void*>(merge_topk_kernel<double, true>)};
    for (int i = 0; i < 4 && err == cudaSuccess; ++i)
      err = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
  });
  ARB_CUDA(err);
  const unsigned grid = static_cast<unsigned>(nq);
  const MergePeers none{};
  if (score_dtype == ARB_F32) {
    const float pad = -3.4028234663852886e38f;
    if (peers)
      merge_topk_kernel<float, true><<<grid, 256, smem, st>>>(nullptr, nullptr, parts, nq, k, n_pow2,
                                                             static_cast<float*>(out_scores), out_ids, pad, *peers);
    else
      merge_topk_kernel<float, false><<<grid, 256, smem, st>>>(static_cast<const float*>(scores), ids, parts, nq, k, n_pow2,
                                                              static_cast<float*>(out_scores), out_ids, pad, none);
  } else {
    if (peers)
      merge_topk_kernel<double, true><<<grid, 256, smem, st>>>(nullptr, nullptr, parts, nq, k, n_pow2,
                                                              static_cast<double*>(out_scores), out_ids, 0.0, *peers);
    else
      merge_topk_kernel<double, false><<<grid, 256, smem, st>>>(static_cast<const double*>(scores), ids, parts, nq, k, n_pow2,
                                                               static_cast<double*>(out_scores), out_ids, 0.0, none);
  }
  ARB_LAUNCH_OK();
  return ARB_OK;
}

int merge_topk_device(int score_dtype, const void* scores, const int64_t* ids, int parts, int64_t nq, int k,
                      void* out_scores, int64_t* out_ids, cudaStream_t st) {
  return merge_topk_launch(score_dtype, scores, ids, nullptr, parts, nq, k, out_scores, out_ids, st);
}


// ----------------------------------------------------------------------------
// Candidate re-scoring of the live pipeline (reference QueryProcessor._hybrid_search,
// query/query_processor.py:3681-3768): per query a list of nc candidates with a vector
// similarity and a BM25 score each, guard flags computed by the host from the note text,
// linear or RRF fusion in fp64 with exactly the reference's operation order, then a stable
// descending sort.  One CTA per query; ranks are counted (O(nc^2 / threads)), nc <= 4096.
// ----------------------------------------------------------------------------
constexpr int RESCORE_MAXC = 4096;
__global__ void __launch_bounds__(256)
rescore_kernel(int method, int nc, const double* __restrict__ vec, const double* __restrict__ bm,
               const int32_t* __restrict__ n_cand, const uint8_t* __restrict__ must_miss,
               const uint8_t* __restrict__ n_ent, const uint8_t* __restrict__ n_pred, double vw, double bw, double rrf_k,
               double* __restrict__ out_score, int32_t* __restrict__ out_order) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  double* sv = reinterpret_cast<double*>(rs_smem);              // [nc] vector similarity
  double* sb = sv + nc;                                         // [nc] bm25 score
  double* sf = sb + nc;                                         // [nc] fused score
  const int64_t q = blockIdx.x;
  const int n = n_cand ? min(nc, max(0, n_cand[q])) : nc;
  const size_t base = static_cast<size_t>(q) * nc;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { sv[i] = vec[base + i]; sb[i] = bm[base + i]; }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double f;
    if (method == ARB_FUSE_LINEAR) {                              // :3709-3737
      double v = sv[i], b = sb[i];
      if (must_miss[base + i]) b = __dmul_rn(b, 0.1);
      for (int e = 0; e < n_ent[base + i]; ++e) v = __dmul_rn(v, 1.2);
      for (int e = 0; e < n_pred[base + i]; ++e) b = __dmul_rn(b, 1.3);
      f = __dadd_rn(__dmul_rn(vw, v), __dmul_rn(bw, b));
    } else {                                                      // :3739-3757: 0-based stable descending ranks
      const double v = sv[i], b = sb[i];
      int rv = 0, rb = 0;
      for (int j = 0; j < n; ++j) {
        const double vj = sv[j], bj = sb[j];
        rv += (vj > v || (vj == v && j < i)) ? 1 : 0;
        rb += (bj > b || (bj == b && j < i)) ? 1 : 0;
      }
      f = __dadd_rn(__ddiv_rn(vw, __dadd_rn(rrf_k, static_cast<double>(rv))),
                    __ddiv_rn(bw, __dadd_rn(rrf_k, static_cast<double>(rb))));
      if (must_miss[base + i]) f = __dmul_rn(f, 0.1);
    }
    sf[i] = f;
    out_score[base + i] = f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nc; i += blockDim.x) {            // :3760 stable sort, descending
    if (i < n) {
      const double f = sf[i];
      int r = 0;
      for (int j = 0; j < n; ++j) { const double fj = sf[j]; r += (fj > f || (fj == f && j < i)) ? 1 : 0; }
      out_order[base + r] = i;
    } else {
      out_order[base + i] = -1;
      out_score[base + i] = 0.0;
    }
  }
}


// ----------------------------------------------------------------------------
// The two other live fusion stages of QueryProcessor, same shape as rescore_kernel (one CTA per query,
// fp64 with the reference's operation order, ranks by counting, stable descending sort):
//   STAGE_V2    _enhanced_hybrid_search_v2 (query/query_processor.py:1089-1143)
//   STAGE_KHOP  _rerank_khop_candidates (:1549-1635) + _rrf_score (:854-861)
// Everything the reference derives from note TEXT (section / lexical multipliers, boosts, coverage,
// consistency) and from hop_no arrives as per-candidate numbers computed by the host; a multiplier the
// reference would skip is passed as 1.0 (x * 1.0 == x bit for bit).
// ----------------------------------------------------------------------------
struct StageParams {
  int nc;
  const int32_t* n_cand;
  const double* a;            // dense similarity
  const double* b;            // sparse / bm25 score
  // V2
  const double* section; const double* lexical; const uint8_t* satisfies; const double* ent_boost; const double* pred_boost;
  double wa, wb, noise;
  // KHOP
  const double* focused; const double* cov; const double* cons; const double* prf_sim; const uint8_t* is_prf;
  const int32_t* group; const double* focused_w; const double* hop_mult;
  double cov_w, cons_w, rrf_lambda; int rrf_k;
  double* out_score; int32_t* out_order; int32_t* out_kept;
};
enum { STAGE_V2 = 0, STAGE_KHOP = 1 };

template <int STAGE>
__global__ void __launch_bounds__(256)
rescore_stage_kernel(const StageParams p) {
  extern __shared__ __align__(16) uint8_t st_smem[];
  const int nc = p.nc;
  double* sa = reinterpret_cast<double*>(st_smem);              // [nc]
  double* sb = sa + nc;                                         // [nc]
  double* sf = sb + nc;                                         // [nc] final score
  double* sp = sf + nc;                                         // [nc] prf similarity (KHOP)
  int* rk = reinterpret_cast<int*>(sp + nc);                    // [3][nc] 1-based ranks: dense, bm25, prf (0 = not prf)
  int* sg = rk + 3 * nc;                                        // [nc] note-id group
  __shared__ int kept;
  const int64_t q = blockIdx.x;
  const int n = p.n_cand ? min(nc, max(0, p.n_cand[q])) : nc;
  const size_t base = static_cast<size_t>(q) * nc;
  if (threadIdx.x == 0) kept = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    sa[i] = p.a[base + i]; sb[i] = p.b[base + i];
    if (STAGE == STAGE_KHOP) {
      const bool prf = p.is_prf && p.is_prf[base + i];
      sp[i] = prf ? p.prf_sim[base + i] : 0.0;
      rk[2 * nc + i] = prf ? 1 : 0;
      sg[i] = p.group ? p.group[base + i] : i;
    }
  }
  __syncthreads();
  if (STAGE == STAGE_KHOP) {
    // 1-based stable descending ranks (:1590-1601): over all candidates for dense / bm25, inside the prf subset for prf
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double v = sa[i], b = sb[i], s = sp[i];
      const bool prf = rk[2 * nc + i] != 0;
      int rv = 1, rb = 1, rp = 1;
      for (int j = 0; j < n; ++j) {
        const double vj = sa[j], bj = sb[j];
        rv += (vj > v || (vj == v && j < i)) ? 1 : 0;
        rb += (bj > b || (bj == b && j < i)) ? 1 : 0;
        if (prf && rk[2 * nc + j] != 0) { const double sj = sp[j]; rp += (sj > s || (sj == s && j < i)) ? 1 : 0; }
      }
      rk[i] = rv; rk[nc + i] = rb;
      sf[i] = prf ? static_cast<double>(rp) : 0.0;              // parked: rk[2] is still read as the prf flag above
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) rk[2 * nc + i] = static_cast<int>(sf[i]);
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double f;
    if (STAGE == STAGE_V2) {
      f = __dadd_rn(__dmul_rn(p.wa, sa[i]), __dmul_rn(p.wb, sb[i]));                     // :1109
      if (p.section) f = __dmul_rn(f, p.section[base + i]);                              // :1112-1114
      if (p.lexical) f = __dmul_rn(f, p.lexical[base + i]);                              // :1117-1119
      const bool ok = !p.satisfies || p.satisfies[base + i];
      if (f < p.noise && !ok) f = 0.0;                                                   // :1122-1123
      if (p.ent_boost) f = __dmul_rn(f, p.ent_boost[base + i]);                          // :1126-1132
      if (p.pred_boost) f = __dmul_rn(f, p.pred_boost[base + i]);
    } else {
      int rv = rk[i], rb = rk[nc + i], rp = rk[2 * nc + i];
      if (p.group) {                        // one RRF entry per note id: the LAST rank written wins (:1593, :858-860)
        const int g = sg[i];
        for (int j = 0; j < n; ++j)
          if (sg[j] == g) { rv = max(rv, rk[j]); rb = max(rb, rk[nc + j]); rp = max(rp, rk[2 * nc + j]); }
      }
      double bonus = __ddiv_rn(1.0, static_cast<double>(p.rrf_k + rv));                  // 0.0 + x == x
      bonus = __dadd_rn(bonus, __ddiv_rn(1.0, static_cast<double>(p.rrf_k + rb)));
      if (rp > 0) bonus = __dadd_rn(bonus, __ddiv_rn(1.0, static_cast<double>(p.rrf_k + rp)));
      f = __dadd_rn(__dmul_rn(p.wa, sa[i]), __dmul_rn(p.wb, sb[i]));                     // base (:1609-1610)
      f = __dadd_rn(f, __dmul_rn(p.focused_w[base + i], p.focused ? p.focused[base + i] : 0.0));
      f = __dadd_rn(f, __dmul_rn(p.cov_w, p.cov ? p.cov[base + i] : 0.0));
      f = __dadd_rn(f, __dmul_rn(p.cons_w, p.cons ? p.cons[base + i] : 0.0));
      f = __dadd_rn(f, __dmul_rn(p.rrf_lambda, bonus));                                  // :1625
      if (p.hop_mult) f = __dmul_rn(f, p.hop_mult[base + i]);                            // :1626-1627
    }
    p.out_score[base + i] = f;
    sf[i] = f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nc; i += blockDim.x) {
    if (i < n) {
      const double f = sf[i];
      if (STAGE == STAGE_V2 && !(f > 0.0)) continue;                                     // :1139
      int r = 0;
      for (int j = 0; j < n; ++j) { const double fj = sf[j]; r += (fj > f || (fj == f && j < i)) ? 1 : 0; }
      p.out_order[base + r] = i;
      if (STAGE == STAGE_V2) atomicAdd(&kept, 1);
    } else {
      p.out_score[base + i] = 0.0;
      if (STAGE == STAGE_KHOP) p.out_order[base + i] = -1;
    }
  }
  if (STAGE == STAGE_V2) {
    __syncthreads();
    for (int i = kept + threadIdx.x; i < nc; i += blockDim.x) p.out_order[base + i] = -1;
    if (threadIdx.x == 0 && p.out_kept) p.out_kept[q] = kept;
  }
}

template <int STAGE>
static int launch_stage(const StageParams& p, int64_t nq, cudaStream_t st) {
  ARB_CHECK(nq >= 0 && p.nc >= 1 && p.nc <= RESCORE_MAXC, ARB_ELIMIT, "rescore: 1 <= nc <= 4096");
  if (nq == 0) return ARB_OK;
  const size_t smem = static_cast<size_t>(p.nc) * (4 * 8 + 4 * 4);
  ARB_CUDA(cudaFuncSetAttribute(rescore_stage_kernel<STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                RESCORE_MAXC * (4 * 8 + 4 * 4)));
  rescore_stage_kernel<STAGE><<<static_cast<unsigned>(nq), 256, smem, st>>>(p);
  ARB_LAUNCH_OK();
  return ARB_OK;
}

}  // namespace arb

using namespace arb;

extern "C" int arb_rescore_v2(int64_t nq, int nc, const double* dense, const double* sparse, const int32_t* n_cand,
                              const double* section_mult, const double* lexical_mult, const uint8_t* satisfies_must,
                              const double* entity_boost, const double* predicate_boost, double dense_weight,
                              double sparse_weight, double noise_threshold, double* out_score, int32_t* out_order,
                              int32_t* out_kept, void* stream) {
  ARB_CHECK(dense && sparse && out_score && out_order, ARB_EINVAL, "arb_rescore_v2: NULL argument");
  StageParams p{};
  p.nc = nc; p.n_cand = n_cand; p.a = dense; p.b = sparse;
  p.section = section_mult; p.lexical = lexical_mult; p.satisfies = satisfies_must;
  p.ent_boost = entity_boost; p.pred_boost = predicate_boost;
  p.wa = dense_weight; p.wb = sparse_weight; p.noise = noise_threshold;
  p.out_score = out_score; p.out_order = out_order; p.out_kept = out_kept;
  return launch_stage<STAGE_V2>(p, nq, static_cast<cudaStream_t>(stream));
}

extern "C" int arb_rescore_khop(int64_t nq, int nc, const double* dense, const double* bm25, const double* focused,
                                const double* cov, const double* cons, const double* prf_sim, const uint8_t* is_prf,
                                const int32_t* group, const double* focused_w, const double* hop_mult,
                                const int32_t* n_cand, double dense_weight, double bm25_weight, double cov_weight,
                                double cons_weight, double rrf_lambda, int rrf_k, double* out_score, int32_t* out_order,
                                void* stream) {
  ARB_CHECK(dense && bm25 && focused_w && out_score && out_order, ARB_EINVAL, "arb_rescore_khop: NULL argument");
  ARB_CHECK(!is_prf || prf_sim, ARB_EINVAL, "arb_rescore_khop: is_prf without prf_sim");
  ARB_CHECK(rrf_k >= 0, ARB_EINVAL, "arb_rescore_khop: rrf_k < 0");
  StageParams p{};
  p.nc = nc; p.n_cand = n_cand; p.a = dense; p.b = bm25;
  p.focused = focused; p.cov = cov; p.cons = cons; p.prf_sim = prf_sim; p.is_prf = is_prf; p.group = group;
  p.focused_w = focused_w; p.hop_mult = hop_mult;
  p.wa = dense_weight; p.wb = bm25_weight; p.cov_w = cov_weight; p.cons_w = cons_weight; p.rrf_lambda = rrf_lambda;
  p.rrf_k = rrf_k;
  p.out_score = out_score; p.out_order = out_order;
  return launch_stage<STAGE_KHOP>(p, nq, static_cast<cudaStream_t>(stream));
}

extern "C" int arb_merge_topk_peers(int score_dtype, const void* const* part_scores, const int64_t* const* part_ids,
                                    int parts, int64_t row0, int64_t nq, int k, void* out_scores, int64_t* out_ids,
                                    void* stream) {
  ARB_CHECK(part_scores && part_ids && out_scores && out_ids, ARB_EINVAL, "arb_merge_topk_peers: NULL argument");
  ARB_CHECK(parts >= 1 && parts <= MERGE_MAX_PEERS, ARB_ELIMIT, "arb_merge_topk_peers: 1 <= parts <= 16");
  ARB_CHECK(row0 >= 0, ARB_EINVAL, "arb_merge_topk_peers: row0 < 0");
  MergePeers p{};
  for (int i = 0; i < parts; ++i) {
    ARB_CHECK(part_scores[i] && part_ids[i], ARB_EINVAL, "arb_merge_topk_peers: NULL list pointer");
    p.scores[i] = part_scores[i];
    p.ids[i] = part_ids[i];
  }
  p.row0 = row0;
  return merge_topk_launch(score_dtype, nullptr, nullptr, &p, parts, nq, k, out_scores, out_ids,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int arb_merge_topk(int score_dtype, const void* scores, const int64_t* ids, int parts, int64_t nq, int k,
                              int positive_only, void* out_scores, int64_t* out_ids, void* stream) {
  (void)positive_only;   // shard lists already exclude non-positive BM25 scores
  ARB_CHECK(scores && ids && out_scores && out_ids, ARB_EINVAL, "arb_merge_topk: NULL argument");
  return merge_topk_device(score_dtype, scores, ids, parts, nq, k, out_scores, out_ids, static_cast<cudaStream_t>(stream));
}

extern "C" int arb_fuse(int method, const double weights[4], double rrf_k, int pool, int64_t nq,
                        const int64_t* const ids[4], const void* const scores[4], const int score_dtype[4],
                        const int k_src[4], int64_t* out_ids, double* out_final, double* out_src,
                        int32_t* out_count, int is_device, void* stream) {
  ARB_CHECK(method == ARB_FUSE_LINEAR || method == ARB_FUSE_RRF, ARB_EINVAL, "arb_fuse: unknown method");
  ARB_CHECK(weights && ids && scores && score_dtype && k_src, ARB_EINVAL, "arb_fuse: NULL argument");
  ARB_CHECK(out_ids && out_final && out_src && out_count, ARB_EINVAL, "arb_fuse: NULL output");
  ARB_CHECK(pool >= 1 && pool <= FUSE_MAXU, ARB_ELIMIT, "arb_fuse: candidate_pool must be in [1, 1024]");
  ARB_CHECK(nq >= 0 && nq <= 2147483647ll, ARB_EINVAL, "arb_fuse: bad nq");
  if (nq == 0) return ARB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FuseParams p;
  memset(&p, 0, sizeof(p));
  p.method = method; p.rrf_k = rrf_k; p.pool = pool; p.nq = nq;
  AsyncBuf din_ids[4], din_sc[4], d_oid, d_ofin, d_osrc, d_ocnt;
  for (int sidx = 0; sidx < 4; ++sidx) {
    p.w[sidx] = weights[sidx];
    p.k[sidx] = (ids[sidx] && scores[sidx]) ? k_src[sidx] : 0;
    ARB_CHECK(p.k[sidx] >= 0 && p.k[sidx] <= FUSE_MAXK, ARB_ELIMIT, "arb_fuse: at most 256 entries per source");
    p.dtype[sidx] = score_dtype[sidx];
    ARB_CHECK(p.k[sidx] == 0 || p.dtype[sidx] == ARB_F32 || p.dtype[sidx] == ARB_F64, ARB_EINVAL, "arb_fuse: score dtype must be f32 or f64");
    p.ids[sidx] = p.k[sidx] ? ids[sidx] : nullptr;
    p.scores[sidx] = p.k[sidx] ? scores[sidx] : nullptr;
    if (!is_device && p.k[sidx]) {
      const size_t n = static_cast<size_t>(nq) * p.k[sidx];
      const size_t esz = p.dtype[sidx] == ARB_F32 ? 4 : 8;
      ARB_CUDA(din_ids[sidx].alloc(n * 8, st));
      ARB_CUDA(din_sc[sidx].alloc(n * esz, st));
      ARB_CUDA(cudaMemcpyAsync(din_ids[sidx].p, ids[sidx], n * 8, cudaMemcpyHostToDevice, st));
      ARB_CUDA(cudaMemcpyAsync(din_sc[sidx].p, scores[sidx], n * esz, cudaMemcpyHostToDevice, st));
      p.ids[sidx] = din_ids[sidx].as<int64_t>();
      p.scores[sidx] = din_sc[sidx].p;
    }
  }
  p.out_ids = out_ids; p.out_final = out_final; p.out_src = out_src; p.out_count = out_count;
  const size_t np = static_cast<size_t>(nq) * pool;
  if (!is_device) {
    ARB_CUDA(d_oid.alloc(np * 8, st)); ARB_CUDA(d_ofin.alloc(np * 8, st));
    ARB_CUDA(d_osrc.alloc(np * 32, st)); ARB_CUDA(d_ocnt.alloc(static_cast<size_t>(nq) * 4, st));
    p.out_ids = d_oid.as<int64_t>(); p.out_final = d_ofin.as<double>();
    p.out_src = d_osrc.as<double>(); p.out_count = d_ocnt.as<int32_t>();
  }
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    err = cudaFuncSetAttribute(fuse_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(FuseSmem<128>)));
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(fuse_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(FuseSmem<256>)));
  });
  ARB_CUDA(err);
  {
    ProfileScope ps(PROF_FUSE, st);
    const int kmax = std::max(std::max(p.k[0], p.k[1]), std::max(p.k[2], p.k[3]));
    if (kmax <= 128 && pool <= FuseSmem<128>::MAXU)
      fuse_kernel<128><<<static_cast<unsigned>(nq), FUSE_THREADS, sizeof(FuseSmem<128>), st>>>(p);
    else
      fuse_kernel<256><<<static_cast<unsigned>(nq), FUSE_THREADS, sizeof(FuseSmem<256>), st>>>(p);
  }
  ARB_LAUNCH_OK();
  if (!is_device) {
    ARB_CUDA(cudaMemcpyAsync(out_ids, p.out_ids, np * 8, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaMemcpyAsync(out_final, p.out_final, np * 8, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaMemcpyAsync(out_src, p.out_src, np * 32, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaMemcpyAsync(out_count, p.out_count, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaStreamSynchronize(st));
  }
  return ARB_OK;
}

extern "C" int arb_rescore_candidates(int method, int64_t nq, int nc, const double* vec, const double* bm25,
                                      const int32_t* n_cand, const uint8_t* must_miss, const uint8_t* n_ent,
                                      const uint8_t* n_pred, double vector_weight, double bm25_weight, double rrf_k,
                                      double* out_score, int32_t* out_order, void* stream) {
  ARB_CHECK(method == ARB_FUSE_LINEAR || method == ARB_FUSE_RRF, ARB_EINVAL, "arb_rescore_candidates: unknown method");
  ARB_CHECK(nq >= 0 && nc >= 0 && nc <= RESCORE_MAXC, ARB_ELIMIT, "arb_rescore_candidates: at most 4096 candidates per query");
  if (nq == 0 || nc == 0) return ARB_OK;
  ARB_CHECK(vec && bm25 && must_miss && n_ent && n_pred && out_score && out_order, ARB_EINVAL, "arb_rescore_candidates: NULL argument");
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    err = cudaFuncSetAttribute(rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RESCORE_MAXC * 24);
  });
  ARB_CUDA(err);
  rescore_kernel<<<static_cast<unsigned>(nq), 256, static_cast<size_t>(nc) * 24, static_cast<cudaStream_t>(stream)>>>(
      method, nc, vec, bm25, n_cand, must_miss, n_ent, n_pred, vector_weight, bm25_weight, rrf_k, out_score, out_order);
  ARB_LAUNCH_OK();
  return ARB_OK;
}
