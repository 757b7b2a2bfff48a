// BM25 scoring over a block-partitioned inverted index (sm_100a).
//
// Replaces corpus.get_scores() at reference utils/bm25_search.py:309 (rank_bm25
// BM25Okapi) / :43-63 (SimpleBM25) and the max-normalisation :329-333.
//
// Bit-exactness: the reference adds, per query token in order,
//     idf[t] * ( tf*(k1+1) / (tf + k1*(1 - b + b*dl/avgdl)) )          (fp64)
// to every doc.  The parenthesised factor depends only on (tf, dl); the host
// tabulates it once per distinct pair with the reference's own numpy expression
// (`gtab`), a posting stores the table index, and the kernel performs exactly the
// two remaining roundings: one fp64 multiply and one fp64 add, token by token in
// query order (a block barrier separates tokens, so no two tokens race on a doc).
//
// Layout in HBM: postings are term-major, doc-sorted, and cut at doc-block
// boundaries (DOC_BLOCK = 16384 docs = one 128 KB fp64 accumulator tile in shared
// memory).  post[i] = gcode << 14 | (doc % DOC_BLOCK)  (4 B per posting);
// blk_off[t*NB + b] is where block b of term t starts.
#include "bm25.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace arb {

constexpr int BM25_THREADS = 1024;
constexpr int CAND_CAP = 2048;                        // candidate buffer (power of two)
constexpr int GTAB_SMEM = 4096;                       // g-table entries cached in smem
constexpr int BM25_SMEM = DOC_BLOCK * 8 + CAND_CAP * 12 + GTAB_SMEM * 8 + 2 * ARB_MAX_QUERY_TERMS * 4 + 64;

// ---------------------------------------------------------------- build kernels
__global__ void bm25_pack_kernel(const int32_t* __restrict__ doc, const int32_t* __restrict__ gcode, int64_t n,
                                 uint32_t* __restrict__ post) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) post[i] = (static_cast<uint32_t>(gcode[i]) & GCODE_MASK) << DOC_BLOCK_LOG2 | (static_cast<uint32_t>(doc[i]) & (DOC_BLOCK - 1));
}

// one thread per (term, block): first posting of term t whose doc >= b*DOC_BLOCK
__global__ void bm25_blkoff_kernel(const int64_t* __restrict__ term_ptr, const int32_t* __restrict__ doc,
                                   int64_t n_terms, int nb, int64_t n_post, uint32_t* __restrict__ blk_off) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i == n_terms * nb) { blk_off[i] = static_cast<uint32_t>(n_post); return; }
  if (i > n_terms * nb) return;
  const int64_t t = i / nb;
  const int32_t lo_doc = static_cast<int32_t>(i % nb) << DOC_BLOCK_LOG2;
  int64_t lo = term_ptr[t], hi = term_ptr[t + 1];
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (doc[mid] < lo_doc) lo = mid + 1; else hi = mid;
  }
  blk_off[i] = static_cast<uint32_t>(lo);
}

// one thread per (dense term, slice boundary): first posting of the term whose doc >= slice*512
__global__ void bm25_fineoff_kernel(const int64_t* __restrict__ term_ptr, const int32_t* __restrict__ doc,
                                    const int32_t* __restrict__ dense_terms, int n_dense, int n_slices,
                                    uint32_t* __restrict__ fine_off) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<int64_t>(n_dense) * (n_slices + 1)) return;
  const int t = dense_terms[i / (n_slices + 1)];
  const int64_t sl = i % (n_slices + 1);
  int64_t lo = term_ptr[t], hi = term_ptr[t + 1];
  const int64_t lo_doc = sl << (DOC_BLOCK_LOG2 - 5);
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (doc[mid] < lo_doc) lo = mid + 1; else hi = mid;
  }
  fine_off[i] = static_cast<uint32_t>(lo);
}

// one thread per posting of a plane term: planes[tile][slot][doc % 256] = idf * g  (the term's whole
// contribution: the same single rounding the per-posting path performs)
__global__ void bm25_plane_fill_kernel(const int32_t* __restrict__ doc, const int32_t* __restrict__ gcode, int64_t lo,
                                       int64_t hi, double idf, const double* __restrict__ gtab, int slot, int n_planes,
                                       double* __restrict__ planes) {
  const int64_t i = lo + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= hi) return;
  const int64_t d = doc[i];
  planes[((d >> PLANE_TILE_LOG2) * n_planes + slot) * PLANE_TILE + plane_pos(static_cast<uint32_t>(d) & (PLANE_TILE - 1))] = __dmul_rn(idf, gtab[gcode[i]]);
}

// per-tile and global maxima of every plane (upper bounds the plane pass prunes with); absent docs count as +0.0
__global__ void bm25_plane_tmax_kernel(const double* __restrict__ planes, int64_t n_tiles, int n_planes,
                                       double* __restrict__ tmax) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n_tiles * n_planes) return;
  const double* row = planes + i * PLANE_TILE;
  double m = 0.0;
  for (int d = 0; d < PLANE_TILE; ++d) m = fmax(m, row[d]);
  tmax[i] = m;
}
__global__ void bm25_plane_gmax_kernel(const double* __restrict__ tmax, int64_t n_tiles, int n_planes, double* __restrict__ gmax) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_planes) return;
  double m = 0.0;
  for (int64_t t = 0; t < n_tiles; ++t) m = fmax(m, tmax[t * n_planes + s]);
  gmax[s] = m;
}

// fp32 copy of the planes for the screening pass, rounded up (so fp32 sums with upward rounding bound the fp64 sums)
__global__ void bm25_plane32_kernel(const double* __restrict__ planes, int64_t n_rows, float* __restrict__ planes32) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n_rows * PLANE_TILE) return;
  const int64_t row = i >> PLANE_TILE_LOG2;
  const uint32_t d = static_cast<uint32_t>(i) & (PLANE_TILE - 1);
  planes32[row * PLANE_TILE + pos32(d)] = __double2float_ru(planes[row * PLANE_TILE + plane_pos(d)]);
}

// term owning posting i: the last t with term_ptr[t] <= i
__device__ __forceinline__ int64_t term_of_posting(const int64_t* __restrict__ term_ptr, int64_t n_terms, int64_t i) {
  int64_t lo = 0, hi = n_terms;                                  // answer in [lo, hi)
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (term_ptr[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}
// doc-major rows, step 1: sort key of every posting = its doc (sparse term with idf != 0) or the sentinel n_docs
__global__ void bm25_sparse_key_kernel(const int64_t* __restrict__ term_ptr, int64_t n_terms, const int32_t* __restrict__ doc,
                                       const int* __restrict__ plane_slot, const double* __restrict__ idf, int64_t n_post,
                                       uint32_t sentinel, uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n_post) return;
  const int64_t t = term_of_posting(term_ptr, n_terms, i);
  key[i] = (plane_slot[t] >= 0 || idf[t] == 0.0) ? sentinel : static_cast<uint32_t>(doc[i]);
  val[i] = static_cast<uint32_t>(i);
}
// step 2 (after a stable sort by key, which keeps a doc's postings in ascending term order): row offsets
__global__ void bm25_drow_off_kernel(const uint32_t* __restrict__ sorted_key, int64_t n_post, int64_t n_docs,
                                     uint32_t* __restrict__ drow_off) {
  const int64_t d = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (d > n_docs) return;
  int64_t lo = 0, hi = n_post;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (sorted_key[mid] < static_cast<uint32_t>(d)) lo = mid + 1; else hi = mid;
  }
  drow_off[d] = static_cast<uint32_t>(lo);
}
// tile-major copy: key = tile of the posting's doc (sentinel n_tiles for plane / zero-idf terms); the stable sort keeps
// the (term, doc) order of the term-major input inside a tile
__global__ void bm25_tile_key_kernel(const uint32_t* __restrict__ doc_key, int64_t n_post, uint32_t doc_sentinel,
                                     uint32_t tile_sentinel, uint32_t* __restrict__ key) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n_post) return;
  const uint32_t d = doc_key[i];
  key[i] = d == doc_sentinel ? tile_sentinel : d >> PLANE_TILE_LOG2;
}
__global__ void bm25_tt_fill_kernel(const uint32_t* __restrict__ sorted_val, int64_t n_sp, const int64_t* __restrict__ term_ptr,
                                    int64_t n_terms, const int32_t* __restrict__ doc, const int32_t* __restrict__ gcode,
                                    const double* __restrict__ idf, const double* __restrict__ gtab, uint32_t* __restrict__ tt_term,
                                    uint8_t* __restrict__ tt_doc, double* __restrict__ tt_val, uint32_t* __restrict__ tt_bloom) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n_sp) return;
  const int64_t p = sorted_val[i];
  const int64_t t = term_of_posting(term_ptr, n_terms, p);
  tt_term[i] = static_cast<uint32_t>(t);
  tt_doc[i] = static_cast<uint8_t>(doc[p] & (PLANE_TILE - 1));
  tt_val[i] = __dmul_rn(idf[t], gtab[gcode[p]]);
  atomicOr(tt_bloom + static_cast<int64_t>(doc[p]) * BLOOM_W + bloom_word(static_cast<uint32_t>(t)), bloom_mask(static_cast<uint32_t>(t)));
}
// run structure of the tile rows, static per index: the member pass joins whole runs (the postings of one term inside a
// tile) with the batch's records and no longer rediscovers heads and lengths for every batch.  One block per tile.
__global__ void bm25_tt_info_kernel(const uint32_t* __restrict__ tt_off, const uint32_t* __restrict__ tt_term,
                                    const uint8_t* __restrict__ tt_doc, uint32_t* __restrict__ tt_info) {
  const uint32_t lo = tt_off[blockIdx.x], hi = tt_off[blockIdx.x + 1];
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const uint32_t t = tt_term[i];
    const uint32_t head = (i == lo || tt_term[i - 1] != t) ? 1u : 0u;
    uint32_t rem = 1u;                                          // a term occurs once per doc: a run holds at most 256 postings
    while (i + rem < hi && tt_term[i + rem] == t) ++rem;
    tt_info[i] = static_cast<uint32_t>(tt_doc[i]) | ((rem - 1u) << 8) | (head << 31);
  }
}
// step 3: term id and finished contribution idf * g of every row entry
__global__ void bm25_drow_fill_kernel(const uint32_t* __restrict__ sorted_val, int64_t n_sp, const int64_t* __restrict__ term_ptr,
                                      int64_t n_terms, const int32_t* __restrict__ gcode, const double* __restrict__ idf,
                                      const double* __restrict__ gtab, uint32_t* __restrict__ drow_term,
                                      double* __restrict__ drow_val) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n_sp) return;
  const int64_t p = sorted_val[i];
  const int64_t t = term_of_posting(term_ptr, n_terms, p);
  drow_term[i] = static_cast<uint32_t>(t);
  drow_val[i] = __dmul_rn(idf[t], gtab[gcode[p]]);
}

// ---------------------------------------------------------------- scoring core
struct Bm25Smem {
  double* acc;                  // [DOC_BLOCK]
  unsigned long long* ckey;     // [CAND_CAP]
  uint32_t* cdoc;               // [CAND_CAP]
  double* gtab;                 // [GTAB_SMEM]
  uint32_t* range;              // [2*MAX_QUERY_TERMS]
  int* misc;                    // [0] candidate count
};

__device__ __forceinline__ Bm25Smem carve(uint8_t* base) {
  Bm25Smem s;
  s.acc = reinterpret_cast<double*>(base);
  s.ckey = reinterpret_cast<unsigned long long*>(base + DOC_BLOCK * 8);
  s.gtab = reinterpret_cast<double*>(base + DOC_BLOCK * 8 + CAND_CAP * 8);
  s.cdoc = reinterpret_cast<uint32_t*>(base + DOC_BLOCK * 8 + CAND_CAP * 8 + GTAB_SMEM * 8);
  s.range = s.cdoc + CAND_CAP;
  s.misc = reinterpret_cast<int*>(s.range + 2 * ARB_MAX_QUERY_TERMS);
  return s;
}

// Accumulate the scores of one doc block for one query into s.acc (fp64, query
// token order).  All threads of the CTA call this; ends with a barrier.
__device__ __forceinline__ void bm25_accumulate_block(const Bm25Dev& ix, const Bm25Smem& s, const int32_t* q_terms,
                                                      int max_qt, int b, bool g_in_smem, const double* init = nullptr,
                                                      int n_init = 0) {
  const int tid = threadIdx.x;
  // `init`: the sums so far of a query longer than one call's token limit (the chunks continue the same in-order sum)
  for (int i = tid; i < DOC_BLOCK; i += BM25_THREADS) s.acc[i] = (init && i < n_init) ? init[i] : 0.0;
  if (tid < max_qt) {
    const int t = q_terms[tid];
    uint32_t lo = 0, hi = 0;
    if (t >= 0 && t < ix.n_terms) {
      lo = __ldg(ix.blk_off + static_cast<int64_t>(t) * ix.nb + b);
      hi = __ldg(ix.blk_off + static_cast<int64_t>(t) * ix.nb + b + 1);
    }
    s.range[2 * tid] = lo;
    s.range[2 * tid + 1] = hi;
  }
  __syncthreads();
  for (int j = 0; j < max_qt; ++j) {
    const uint32_t lo = s.range[2 * j], hi = s.range[2 * j + 1];
    if (lo == hi) continue;                                   // uniform across the CTA
    const double idf = __ldg(ix.idf + q_terms[j]);
    if (idf != 0.0) {                                          // (idf.get(q) or 0): adds +0.0, a no-op
      for (uint32_t i = lo + tid; i < hi; i += BM25_THREADS) {
        const uint32_t u = __ldg(ix.post + i);
        const uint32_t code = post_code(u);
        const double g = g_in_smem ? s.gtab[code] : __ldg(ix.gtab + code);
        const uint32_t d = post_doc(u);
        s.acc[d] = __dadd_rn(s.acc[d], __dmul_rn(idf, g));
      }
    }
    __syncthreads();
  }
}

// block-wide bitonic sort of (ckey desc, cdoc asc) over CAND_CAP slots
__device__ __forceinline__ void bitonic_sort_cands(const Bm25Smem& s) {
  const int tid = threadIdx.x;
  for (int size = 2; size <= CAND_CAP; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = tid; i < CAND_CAP / 2; i += BM25_THREADS) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc_block = ((lo & size) == 0);            // first-half blocks sort "best first"
        const unsigned long long ka = s.ckey[lo], kb = s.ckey[hi];
        const uint32_t da = s.cdoc[lo], db = s.cdoc[hi];
        const bool a_before_b = (ka > kb) || (ka == kb && da < db);
        if (a_before_b != desc_block) {
          s.ckey[lo] = kb; s.ckey[hi] = ka;
          s.cdoc[lo] = db; s.cdoc[hi] = da;
        }
      }
    }
  }
  __syncthreads();
}

// ----------------------------------------------------------------------------
// Top-k kernel.  grid (nq, n_split); one CTA of 16 warps walks its doc blocks.
//
// Ownership: warp w (of 32) owns docs [512*w, 512*w+512) of the current 16K-doc block, i.e.
// its own 4 KB slice of the fp64 accumulator tile.  Only that warp touches the slice,
// so the query-token order of the per-doc additions needs nothing stronger than
// __syncwarp() between terms (the first version used ~500 block barriers per query
// and was bound by shared-memory bank-conflict replays; profiles/r1_ncu_summary.md).
//   * dense terms (df >= n_docs/256) carry per-slice offsets (fine_off): the warp
//     streams exactly its postings from L2 with unrolled, coalesced 4-byte loads;
//   * sparse terms have few postings per block: every warp scans the block segment
//     and keeps the postings of its own slice.
// The (tf, dl) factor table is replicated across shared-memory banks so the random
// table lookups of a warp do not serialise.  After the last term each warp scans its
// slice (fused with re-zeroing) and appends docs that beat the running k-th
// (score, doc) pair to the CTA's candidate buffer.
// ----------------------------------------------------------------------------
constexpr int TOPK_WARPS = 32;
constexpr int TOPK_THREADS = TOPK_WARPS * 32;
constexpr int CAP2 = 1024;                           // candidate buffer (power of two)
constexpr int GREP_BYTES = 65536;                    // replicated factor table
constexpr int TOPK_SMEM = DOC_BLOCK * 8 + 256 + GREP_BYTES + CAP2 * 12 + ARB_MAX_QUERY_TERMS * 16 + 64;
constexpr uint32_t NULL_POSTING = 0xFFFFFFFFu;       // never a real posting: n_g < 2^18

struct TopkSmem {
  double* acc; double* grep; unsigned long long* ckey; uint32_t* cdoc; double* idf; int* dslot; int* misc;
  const double* gglobal;
};
__device__ __forceinline__ TopkSmem carve_topk(uint8_t* b) {
  TopkSmem s;
  s.acc = reinterpret_cast<double*>(b);                 b += DOC_BLOCK * 8 + 256;   // + one dummy slot per lane
  s.grep = reinterpret_cast<double*>(b);                b += GREP_BYTES;
  s.ckey = reinterpret_cast<unsigned long long*>(b);    b += CAP2 * 8;
  s.idf = reinterpret_cast<double*>(b);                 b += ARB_MAX_QUERY_TERMS * 8;
  s.cdoc = reinterpret_cast<uint32_t*>(b);              b += CAP2 * 4;
  s.dslot = reinterpret_cast<int*>(b);                  b += ARB_MAX_QUERY_TERMS * 4;
  s.misc = reinterpret_cast<int*>(b);
  s.gglobal = nullptr;
  return s;
}

// bitonic sort of the candidate buffer by (key desc, doc asc); whole CTA
__device__ __forceinline__ void sort_cands(const TopkSmem& s, int tid) {
  for (int size = 2; size <= CAP2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = tid; i < CAP2 / 2; i += TOPK_THREADS) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool best_first = ((lo & size) == 0);
        const unsigned long long ka = s.ckey[lo], kb = s.ckey[hi];
        const uint32_t da = s.cdoc[lo], db = s.cdoc[hi];
        const bool a_before_b = (ka > kb) || (ka == kb && da < db);
        if (a_before_b != best_first) { s.ckey[lo] = kb; s.ckey[hi] = ka; s.cdoc[lo] = db; s.cdoc[hi] = da; }
      }
    }
  }
  __syncthreads();
}

// one posting -> acc[doc] += idf * g[code]; a NULL posting goes to the lane's dummy slot behind the tile
template <int REP_LOG2, bool CHECK>
__device__ __forceinline__ void apply_posting(const TopkSmem& s, uint32_t u, double idf, int lane) {
  uint32_t code = post_code(u);
  uint32_t doc = post_doc(u);
  if (CHECK) {
    const bool null = u == NULL_POSTING;
    code = null ? 0u : code;
    doc = null ? static_cast<uint32_t>(DOC_BLOCK + lane) : doc;
  }
  double g;
  if (REP_LOG2 < 0) g = __ldg(s.gglobal + code);
  else g = s.grep[(code << (REP_LOG2 < 0 ? 0 : REP_LOG2)) + (lane & ((1 << (REP_LOG2 < 0 ? 0 : REP_LOG2)) - 1))];
  s.acc[doc] = __dadd_rn(s.acc[doc], __dmul_rn(idf, g));
}

// Dense term: every posting in [lo, hi) belongs to this warp's slice (hi - lo <= SLICE).
// Register double-buffered: the next 128 postings are in flight while the current 128 are applied.
template <int REP_LOG2>
__device__ __forceinline__ void accumulate_dense(const TopkSmem& s, const uint32_t* __restrict__ post, uint32_t lo,
                                                 uint32_t hi, double idf, int lane) {
  constexpr int U = 4;
  uint32_t cur[U], nxt[U];
  uint32_t base = lo;
#pragma unroll
  for (int r = 0; r < U; ++r) { const uint32_t i = base + r * 32 + lane; cur[r] = i < hi ? __ldg(post + i) : NULL_POSTING; }
  while (true) {
    const uint32_t nbase = base + 32 * U;
    const bool more = nbase < hi;
    if (more) {
#pragma unroll
      for (int r = 0; r < U; ++r) { const uint32_t i = nbase + r * 32 + lane; nxt[r] = i < hi ? __ldg(post + i) : NULL_POSTING; }
    }
    if (nbase <= hi) {                               // full chunk: no NULL postings
#pragma unroll
      for (int r = 0; r < U; ++r) apply_posting<REP_LOG2, false>(s, cur[r], idf, lane);
    } else {
#pragma unroll
      for (int r = 0; r < U; ++r) apply_posting<REP_LOG2, true>(s, cur[r], idf, lane);
    }
    if (!more) break;
#pragma unroll
    for (int r = 0; r < U; ++r) cur[r] = nxt[r];
    base = nbase;
  }
}

// Sparse term: few postings per 16K block; every warp scans them and keeps those of its own slice.
template <int REP_LOG2>
__device__ __forceinline__ void accumulate_sparse(const TopkSmem& s, const uint32_t* __restrict__ post, uint32_t lo,
                                                  uint32_t hi, double idf, int lane, int warp) {
  for (uint32_t i = lo + lane; i < hi; i += 32) {
    const uint32_t u = __ldg(post + i);
    if (static_cast<int>(post_doc(u) >> SLICE_LOG2) == warp) apply_posting<REP_LOG2, false>(s, u, idf, lane);
  }
}

template <int REP_LOG2>
__device__ __forceinline__ void bm25_topk_body(const Bm25Dev& ix, const TopkSmem& s, const int32_t* __restrict__ q_terms,
                                               int max_qt, int k, int64_t id_base, int b0, int b1,
                                               double* __restrict__ os, int64_t* __restrict__ oi) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the (up to two) query terms whose offsets this lane fetches for the warp
  int t_a = -1, t_b = -1, d_a = -1, d_b = -1;
  if (lane < max_qt) { t_a = q_terms[lane]; if (t_a >= ix.n_terms) t_a = -1; }
  if (lane + 32 < max_qt) { t_b = q_terms[lane + 32]; if (t_b >= ix.n_terms) t_b = -1; }
  if (t_a >= 0 && s.idf[lane] == 0.0) t_a = -1;            // idf 0 adds +0.0: a no-op in the reference
  if (t_b >= 0 && s.idf[lane + 32] == 0.0) t_b = -1;
  if (t_a >= 0) d_a = s.dslot[lane];
  if (t_b >= 0) d_b = s.dslot[lane + 32];
  const int nsl = ix.n_slices;

  auto load_range = [&](int t, int dslot, int b, uint32_t& lo, uint32_t& hi) {
    lo = hi = 0;
    if (t < 0 || b >= b1) return;
    if (dslot >= 0) {
      const int64_t sl = (static_cast<int64_t>(b) << (DOC_BLOCK_LOG2 - SLICE_LOG2)) + warp;
      if (sl < nsl) {
        lo = __ldg(ix.fine_off + static_cast<int64_t>(dslot) * (nsl + 1) + sl);
        hi = __ldg(ix.fine_off + static_cast<int64_t>(dslot) * (nsl + 1) + sl + 1);
      }
    } else {
      lo = __ldg(ix.blk_off + static_cast<int64_t>(t) * ix.nb + b);
      hi = __ldg(ix.blk_off + static_cast<int64_t>(t) * ix.nb + b + 1);
    }
  };

  unsigned long long thr_key = f64_key(0.0);     // a candidate needs score > 0 ...
  uint32_t thr_doc = 0u;                          // ... or a tie with the k-th score at a smaller doc id
  double2* my_acc2 = reinterpret_cast<double2*>(s.acc + (warp << SLICE_LOG2));

  uint32_t lo_a, hi_a, lo_b, hi_b;
  load_range(t_a, d_a, b0, lo_a, hi_a);
  load_range(t_b, d_b, b0, lo_b, hi_b);
  for (int b = b0; b < b1; ++b) {
    uint32_t nlo_a, nhi_a, nlo_b, nhi_b;          // prefetch the next block's ranges
    load_range(t_a, d_a, b + 1, nlo_a, nhi_a);
    load_range(t_b, d_b, b + 1, nlo_b, nhi_b);
    // which terms have postings for this warp in this block (bit j of the masks)
    const uint32_t ne_a = __ballot_sync(0xffffffffu, hi_a > lo_a);
    const uint32_t ne_b = __ballot_sync(0xffffffffu, hi_b > lo_b);
    const uint32_t dn_a = __ballot_sync(0xffffffffu, d_a >= 0);
    const uint32_t dn_b = __ballot_sync(0xffffffffu, d_b >= 0);
    const bool touched = (ne_a | ne_b) != 0u;
    for (int half = 0; half < 2; ++half) {
      uint32_t ne = half ? ne_b : ne_a;
      const uint32_t dn = half ? dn_b : dn_a;
      while (ne) {                                 // ascending j = query token order
        const int jl = __ffs(ne) - 1;
        ne &= ne - 1;
        const uint32_t lo = __shfl_sync(0xffffffffu, half ? lo_b : lo_a, jl);
        const uint32_t hi = __shfl_sync(0xffffffffu, half ? hi_b : hi_a, jl);
        const double idf = s.idf[half * 32 + jl];
        if ((dn >> jl) & 1u) accumulate_dense<REP_LOG2>(s, ix.post, lo, hi, idf, lane);
        else                 accumulate_sparse<REP_LOG2>(s, ix.post, lo, hi, idf, lane, warp);
        __syncwarp();                              // the next term may hit the same docs from other lanes
      }
    }
    lo_a = nlo_a; hi_a = nhi_a; lo_b = nlo_b; hi_b = nhi_b;

    // ---- harvest this warp's slice and re-zero it; re-scan after a compaction if the buffer overflowed
    const uint32_t doc0 = (static_cast<uint32_t>(b) << DOC_BLOCK_LOG2) + (static_cast<uint32_t>(warp) << SLICE_LOG2);
    bool need_scan = touched;
    while (true) {
      if (need_scan) {
#pragma unroll 2
        for (int i = 0; i < SLICE / 64; ++i) {
          const double2 v2 = my_acc2[i * 32 + lane];
          if (v2.x != 0.0 || v2.y != 0.0) {
            double2 left = make_double2(0.0, 0.0);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const double v = h ? v2.y : v2.x;
              if (v != 0.0) {
                const uint32_t doc = doc0 + (i * 32 + lane) * 2 + h;
                const unsigned long long key = f64_key(v);
                if (key > thr_key || (key == thr_key && doc < thr_doc)) {
                  const int pos = atomicAdd(&s.misc[0], 1);
                  if (pos < CAP2) { s.ckey[pos] = key; s.cdoc[pos] = doc; }
                  else { s.misc[1] = 1; if (h) left.y = v; else left.x = v; }   // overflowed: leave for the re-scan
                }
              }
            }
            my_acc2[i * 32 + lane] = left;
          }
        }
      }
      __syncthreads();
      const int overflow = s.misc[1];
      const int cnt = min(s.misc[0], CAP2);
      __syncthreads();                             // everyone has read the counters before anyone moves on
      if (!overflow && cnt <= CAP2 / 2) break;
      // buffer filling up: keep the k best, tighten the threshold, re-scan what was left behind
      if (tid == 0) s.misc[0] = cnt;
      sort_cands(s, tid);
      const int keep = min(cnt, k);
      if (cnt >= k) { thr_key = s.ckey[k - 1]; thr_doc = s.cdoc[k - 1]; }
      __syncthreads();
      for (int i = keep + tid; i < CAP2; i += TOPK_THREADS) { s.ckey[i] = 0ull; s.cdoc[i] = 0xFFFFFFFFu; }
      if (tid == 0) { s.misc[0] = keep; s.misc[1] = 0; }
      __syncthreads();
      if (!overflow) break;
      need_scan = true;
    }
  }
  // ---- final ordering and output
  __syncthreads();
  sort_cands(s, tid);
  const int cnt = min(min(s.misc[0], CAP2), k);
  for (int i = tid; i < k; i += TOPK_THREADS) {
    if (i < cnt) { os[i] = key_f64(s.ckey[i]); oi[i] = id_base + static_cast<int64_t>(s.cdoc[i]); }
    else         { os[i] = 0.0; oi[i] = -1; }
  }
}

__global__ void __launch_bounds__(TOPK_THREADS, 1)
bm25_topk_kernel(const Bm25Dev ix, const int32_t* __restrict__ q_terms_all, int max_qt, int k, int64_t id_base,
                 int blocks_per_split, int64_t nq, const int* __restrict__ flag_list,
                 const int* __restrict__ flag_count, double* __restrict__ part_scores,
                 int64_t* __restrict__ part_ids) {
  extern __shared__ __align__(128) uint8_t bsm[];
  TopkSmem s = carve_topk(bsm);
  s.gglobal = ix.gtab;
  const int tid = threadIdx.x;
  const int split = blockIdx.y;
  const int b0 = split * blocks_per_split;
  const int b1 = min(b0 + blocks_per_split, ix.nb);
  // replication factor of the factor table: as many bank-shifted copies as fit in 64 KB (at most 8)
  int rep_log2 = -1;
  if (ix.n_g * 8 <= GREP_BYTES) {
    rep_log2 = 0;
    while (rep_log2 < 3 && ((ix.n_g * 8) << (rep_log2 + 1)) <= GREP_BYTES) ++rep_log2;
  }
  if (rep_log2 >= 0) {
    const int rep = 1 << rep_log2;
    for (int i = tid; i < ix.n_g * rep; i += TOPK_THREADS) s.grep[i] = ix.gtab[i >> rep_log2];
  }
  // Without a flag list CTA x scores query x; with one (the batched pipeline's fallback) CTA x walks
  // the listed queries x, x + gridDim.x, ...
  const int n_listed = flag_list ? *flag_count : 1;
  for (int slot = flag_list ? blockIdx.x : 0; slot < n_listed; slot += gridDim.x) {
    const int64_t q = flag_list ? flag_list[slot] : blockIdx.x;
    const int32_t* q_terms = q_terms_all + q * max_qt;
    __syncthreads();                                 // the previous query's output loop is done with the buffers
    for (int i = tid; i < DOC_BLOCK / 2; i += TOPK_THREADS) reinterpret_cast<double2*>(s.acc)[i] = make_double2(0.0, 0.0);
    for (int i = tid; i < CAP2; i += TOPK_THREADS) { s.ckey[i] = 0ull; s.cdoc[i] = 0xFFFFFFFFu; }
    if (tid < max_qt) {
      const int t = q_terms[tid];
      const bool ok = t >= 0 && t < ix.n_terms;
      s.idf[tid] = ok ? ix.idf[t] : 0.0;
      s.dslot[tid] = ok ? ix.dense_slot[t] : -1;
    }
    if (tid == 0) { s.misc[0] = 0; s.misc[1] = 0; }
    __syncthreads();
    double* os = part_scores + (static_cast<int64_t>(split) * nq + q) * k;
    int64_t* oi = part_ids + (static_cast<int64_t>(split) * nq + q) * k;
    switch (rep_log2) {
      case 3:  bm25_topk_body<3>(ix, s, q_terms, max_qt, k, id_base, b0, b1, os, oi); break;
      case 2:  bm25_topk_body<2>(ix, s, q_terms, max_qt, k, id_base, b0, b1, os, oi); break;
      case 1:  bm25_topk_body<1>(ix, s, q_terms, max_qt, k, id_base, b0, b1, os, oi); break;
      case 0:  bm25_topk_body<0>(ix, s, q_terms, max_qt, k, id_base, b0, b1, os, oi); break;
      default: bm25_topk_body<-1>(ix, s, q_terms, max_qt, k, id_base, b0, b1, os, oi); break;
    }
  }
}

// Full score vectors: grid (nq, nb); out[q, doc] for doc < n_out.
__global__ void __launch_bounds__(BM25_THREADS, 1)
bm25_scores_kernel(const Bm25Dev ix, const int32_t* __restrict__ q_terms_all, int max_qt, int64_t n_out,
                   double* __restrict__ out, int accumulate) {
  extern __shared__ __align__(16) uint8_t bsm[];
  const Bm25Smem s = carve(bsm);
  const int tid = threadIdx.x;
  const int64_t q = blockIdx.x;
  const int b = blockIdx.y;
  const bool g_in_smem = ix.n_g <= GTAB_SMEM;
  if (g_in_smem) for (int i = tid; i < ix.n_g; i += BM25_THREADS) s.gtab[i] = ix.gtab[i];
  __syncthreads();
  const int64_t doc0 = static_cast<int64_t>(b) << DOC_BLOCK_LOG2;
  const long long lim = min(min(static_cast<long long>(ix.n_docs), static_cast<long long>(n_out)) - static_cast<long long>(doc0), static_cast<long long>(DOC_BLOCK));
  bm25_accumulate_block(ix, s, q_terms_all + q * max_qt, max_qt, b, g_in_smem, accumulate ? out + q * n_out + doc0 : nullptr,
                        static_cast<int>(max(0ll, lim)));
  for (int64_t d = tid; d < lim; d += BM25_THREADS) out[q * n_out + doc0 + d] = s.acc[d];
}

// utils/bm25_search.py:329-333: divide by the max when it is > 0.  One CTA per query.
__global__ void __launch_bounds__(1024)
bm25_normalise_kernel(double* __restrict__ out, int64_t n_out) {
  __shared__ double red[32];
  double* row = out + blockIdx.x * n_out;
  double m = -INFINITY;
  for (int64_t i = threadIdx.x; i < n_out; i += blockDim.x) m = fmax(m, row[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) red[0] = m;
  }
  __syncthreads();
  m = red[0];
  if (m > 0.0)
    for (int64_t i = threadIdx.x; i < n_out; i += blockDim.x) row[i] = __ddiv_rn(row[i], m);
}

struct FieldPtrs { const double* s[8]; double w[8]; };
__global__ void combine_fields_kernel(FieldPtrs f, int n_fields, int64_t n, double* __restrict__ out) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  double total = 0.0;
  for (int k = 0; k < n_fields; ++k) total = __dadd_rn(total, __dmul_rn(f.w[k], f.s[k][i]));
  out[i] = total;
}

__global__ void fill_f64_kernel(double* p, double v, int64_t n) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace arb

using namespace arb;

int arb::launch_bm25_topk(const Bm25Dev& ix, const int32_t* d_terms, int max_qt, int k, int64_t id_base, int64_t nq,
                          int n_split, int blocks_per_split, const int* flag_list, const int* flag_count, int grid_x,
                          double* part_scores, int64_t* part_ids, cudaStream_t st) {
  dim3 grid(static_cast<unsigned>(grid_x), static_cast<unsigned>(n_split));
  bm25_topk_kernel<<<grid, TOPK_THREADS, TOPK_SMEM, st>>>(ix, d_terms, max_qt, k, id_base, blocks_per_split, nq,
                                                          flag_list, flag_count, part_scores, part_ids);
  ARB_LAUNCH_OK();
  return ARB_OK;
}

int arb::bm25_set_smem_attrs() {
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    err = cudaFuncSetAttribute(bm25_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TOPK_SMEM);
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(bm25_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BM25_SMEM);
  });
  ARB_CUDA(err);
  return ARB_OK;
}

extern "C" int arb_bm25_create(arb_bm25** out, int64_t n_docs, int64_t n_terms, const int64_t* term_ptr,
                               const int32_t* post_doc, const int32_t* post_gcode, const double* idf,
                               const double* gtab, int64_t n_g, void* stream) {
  ARB_CHECK(out != nullptr, ARB_EINVAL, "arb_bm25_create: out is NULL");
  ARB_CHECK(n_docs >= 0 && n_docs < (int64_t(1) << 31), ARB_ELIMIT, "arb_bm25_create: n_docs must be < 2^31");
  ARB_CHECK(n_terms >= 0 && term_ptr != nullptr, ARB_EINVAL, "arb_bm25_create: bad term table");
  ARB_CHECK(n_g >= 0 && n_g < (int64_t(1) << GCODE_BITS), ARB_ELIMIT, "arb_bm25_create: 2^18 or more distinct (tf, doc_len) pairs");
  const int64_t n_post = term_ptr[n_terms];
  ARB_CHECK(n_post >= 0 && n_post < 0xFFFFFFF0ll, ARB_ELIMIT, "arb_bm25_create: more than 2^32 postings per shard");
  ARB_CHECK(n_post == 0 || (post_doc && post_gcode && idf && gtab), ARB_EINVAL, "arb_bm25_create: NULL posting arrays");
  const int nb = static_cast<int>(std::max<int64_t>(1, (n_docs + DOC_BLOCK - 1) / DOC_BLOCK));
  ARB_CHECK(n_terms * nb < 0x7FFFFFF0ll, ARB_ELIMIT, "arb_bm25_create: n_terms * doc_blocks too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  arb_bm25* bm = new arb_bm25();
  auto fail = [&](int rc) { arb_bm25_destroy(bm); return rc; };
  const size_t n_off = static_cast<size_t>(n_terms) * nb + 1;
  if (cudaMalloc(&bm->blk_off, n_off * 4) != cudaSuccess || cudaMalloc(&bm->post, (std::max<size_t>(n_post, 1) + 8) * 4) != cudaSuccess ||
      cudaMalloc(&bm->idf, std::max<size_t>(n_terms, 1) * 8) != cudaSuccess || cudaMalloc(&bm->gtab, std::max<size_t>(n_g, 1) * 8) != cudaSuccess) {
    cudaGetLastError();
    set_error("arb_bm25_create: out of device memory");
    return fail(ARB_ENOMEM);
  }
  {
    AsyncBuf d_ptr, d_doc, d_code;
    if (d_ptr.alloc((n_terms + 1) * 8, st) != cudaSuccess || d_doc.alloc(std::max<size_t>(n_post, 1) * 4, st) != cudaSuccess ||
        d_code.alloc(std::max<size_t>(n_post, 1) * 4, st) != cudaSuccess) {
      cudaGetLastError();
      set_error("arb_bm25_create: out of device memory (staging)");
      return fail(ARB_ENOMEM);
    }
    cudaError_t e = cudaMemcpyAsync(d_ptr.p, term_ptr, (n_terms + 1) * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_post) e = cudaMemcpyAsync(d_doc.p, post_doc, n_post * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_post) e = cudaMemcpyAsync(d_code.p, post_gcode, n_post * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_terms) e = cudaMemcpyAsync(bm->idf, idf, n_terms * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_g) e = cudaMemcpyAsync(bm->gtab, gtab, n_g * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail(cuda_fail(e, "bm25 H2D", __FILE__, __LINE__));
    if (n_post) {
      bm25_pack_kernel<<<static_cast<unsigned>((n_post + 255) / 256), 256, 0, st>>>(d_doc.as<int32_t>(), d_code.as<int32_t>(), n_post, bm->post);
      count_launch();
    }
    bm25_blkoff_kernel<<<static_cast<unsigned>((n_off + 255) / 256), 256, 0, st>>>(d_ptr.as<int64_t>(), d_doc.as<int32_t>(), n_terms, nb, n_post, bm->blk_off);
    count_launch();
    // dense terms (df >= n_docs/256, at least 256 postings) get per-512-doc-slice offsets
    {
      const int n_slices = static_cast<int>(std::max<int64_t>(1, (n_docs + 511) / 512));
      std::vector<int32_t> slot(std::max<int64_t>(n_terms, 1), -1), dterms;
      const int64_t thresh = std::max<int64_t>(256, n_docs / 256);
      for (int64_t t = 0; t < n_terms; ++t)
        if (term_ptr[t + 1] - term_ptr[t] >= thresh) { slot[t] = static_cast<int32_t>(dterms.size()); dterms.push_back(static_cast<int32_t>(t)); }
      const size_t n_fine = std::max<size_t>(1, dterms.size() * static_cast<size_t>(n_slices + 1));
      if (cudaMalloc(&bm->dense_slot, slot.size() * 4) != cudaSuccess || cudaMalloc(&bm->fine_off, n_fine * 4) != cudaSuccess) {
        cudaGetLastError(); set_error("arb_bm25_create: out of device memory (fine offsets)"); return fail(ARB_ENOMEM);
      }
      e = cudaMemcpyAsync(bm->dense_slot, slot.data(), slot.size() * 4, cudaMemcpyHostToDevice, st);
      AsyncBuf d_dt;
      if (e == cudaSuccess && !dterms.empty()) {
        e = d_dt.alloc(dterms.size() * 4, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_dt.p, dterms.data(), dterms.size() * 4, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
          bm25_fineoff_kernel<<<static_cast<unsigned>((n_fine + 255) / 256), 256, 0, st>>>(
              d_ptr.as<int64_t>(), d_doc.as<int32_t>(), d_dt.as<int32_t>(), static_cast<int>(dterms.size()), n_slices, bm->fine_off);
          count_launch();
        }
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);       // host vectors die at scope end
      if (e != cudaSuccess) return fail(cuda_fail(e, "bm25 fine offsets", __FILE__, __LINE__));
      bm->dev.n_slices = n_slices;
      bm->dev.n_dense = static_cast<int>(dterms.size());
    }
    // doc-major planes for the densest terms (df >= n_docs/frac, at most MAX_PLANES of them, largest df first)
    {
      int frac = 32, max_planes = MAX_PLANES;
      if (const char* v = std::getenv("ARB_BM25_PLANE_FRAC")) frac = std::max(1, atoi(v));
      if (const char* v = std::getenv("ARB_BM25_MAX_PLANES")) max_planes = std::max(0, std::min(MAX_PLANES, atoi(v)));
      const int64_t min_df = std::max<int64_t>(1, (n_docs + frac - 1) / frac);
      std::vector<std::pair<int64_t, int32_t>> cand;
      for (int64_t t = 0; t < n_terms; ++t) {
        const int64_t f = term_ptr[t + 1] - term_ptr[t];
        if (f >= min_df && idf[t] != 0.0) cand.emplace_back(-f, static_cast<int32_t>(t));
      }
      std::sort(cand.begin(), cand.end());
      if (static_cast<int>(cand.size()) > max_planes) cand.resize(max_planes);
      const int n_planes = static_cast<int>(cand.size());
      const int64_t n_tiles = std::max<int64_t>(1, (n_docs + PLANE_TILE - 1) / PLANE_TILE);
      std::vector<int32_t> pslot(std::max<int64_t>(n_terms, 1), -1);
      for (int s2 = 0; s2 < n_planes; ++s2) pslot[cand[s2].second] = s2;
      const size_t plane_bytes = std::max<size_t>(16, static_cast<size_t>(n_tiles) * n_planes * PLANE_TILE * 8);
      if (cudaMalloc(&bm->plane_slot, pslot.size() * 4) != cudaSuccess || cudaMalloc(&bm->planes, plane_bytes) != cudaSuccess) {
        cudaGetLastError(); set_error("arb_bm25_create: out of device memory (term planes)"); return fail(ARB_ENOMEM);
      }
      e = cudaMemcpyAsync(bm->plane_slot, pslot.data(), pslot.size() * 4, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaMemsetAsync(bm->planes, 0, plane_bytes, st);
      for (int s2 = 0; s2 < n_planes && e == cudaSuccess; ++s2) {
        const int32_t t = cand[s2].second;
        const int64_t lo = term_ptr[t], hi = term_ptr[t + 1];
        bm25_plane_fill_kernel<<<static_cast<unsigned>((hi - lo + 255) / 256), 256, 0, st>>>(
            d_doc.as<int32_t>(), d_code.as<int32_t>(), lo, hi, idf[t], bm->gtab, s2, n_planes, bm->planes);
        count_launch();
      }
      if (e == cudaSuccess && n_planes > 0) {
        const size_t nt = static_cast<size_t>(n_tiles) * n_planes;
        if (cudaMalloc(&bm->plane_tmax, (nt + n_planes) * 8) != cudaSuccess) {
          cudaGetLastError(); set_error("arb_bm25_create: out of device memory (plane maxima)"); return fail(ARB_ENOMEM);
        }
        bm25_plane_tmax_kernel<<<static_cast<unsigned>((nt + 255) / 256), 256, 0, st>>>(bm->planes, n_tiles, n_planes, bm->plane_tmax);
        bm25_plane_gmax_kernel<<<(n_planes + 127) / 128, 128, 0, st>>>(bm->plane_tmax, n_tiles, n_planes, bm->plane_tmax + nt);
        count_launch(2);
      }
      if (e == cudaSuccess && n_planes > 0) {
        const int64_t n_rows = n_tiles * n_planes;
        if (cudaMalloc(&bm->planes32, static_cast<size_t>(n_rows) * PLANE_TILE * 4) != cudaSuccess) {
          cudaGetLastError(); set_error("arb_bm25_create: out of device memory (fp32 planes)"); return fail(ARB_ENOMEM);
        }
        bm25_plane32_kernel<<<static_cast<unsigned>((n_rows * PLANE_TILE + 255) / 256), 256, 0, st>>>(bm->planes, n_rows, bm->planes32);
        count_launch();
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);       // pslot dies at scope end
      if (e != cudaSuccess) return fail(cuda_fail(e, "bm25 term planes", __FILE__, __LINE__));
      bm->dev.n_planes = n_planes;
      bm->dev.n_tiles = n_tiles;
    }
    // doc-major rows of the sparse postings (the batched pipeline joins them with the batch's query-term index)
    {
      int64_t n_sp = 0;
      if (cudaMalloc(&bm->drow_off, static_cast<size_t>(n_docs + 1) * 4) != cudaSuccess) {
        cudaGetLastError(); set_error("arb_bm25_create: out of device memory (doc rows)"); return fail(ARB_ENOMEM);
      }
      if (n_post > 0) {
        AsyncBuf k_in, k_out, v_in, v_out, tmp;
        size_t tmp_bytes = 0;
        const size_t np4 = static_cast<size_t>(n_post) * 4;
        e = k_in.alloc(np4, st);
        if (e == cudaSuccess) e = k_out.alloc(np4, st);
        if (e == cudaSuccess) e = v_in.alloc(np4, st);
        if (e == cudaSuccess) e = v_out.alloc(np4, st);
        int end_bit = 1;
        while ((int64_t(1) << end_bit) <= n_docs) ++end_bit;       // keys are in [0, n_docs]
        if (e == cudaSuccess)
          e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in.as<uint32_t>(), k_out.as<uint32_t>(), v_in.as<uint32_t>(),
                                              v_out.as<uint32_t>(), static_cast<int>(n_post), 0, end_bit, st);
        if (e == cudaSuccess) e = tmp.alloc(tmp_bytes, st);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("arb_bm25_create: out of device memory (doc-row build)"); return fail(ARB_ENOMEM); }
        bm25_sparse_key_kernel<<<static_cast<unsigned>((n_post + 255) / 256), 256, 0, st>>>(
            d_ptr.as<int64_t>(), n_terms, d_doc.as<int32_t>(), bm->plane_slot, bm->idf, n_post, static_cast<uint32_t>(n_docs),
            k_in.as<uint32_t>(), v_in.as<uint32_t>());
        count_launch();
        e = cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, k_in.as<uint32_t>(), k_out.as<uint32_t>(), v_in.as<uint32_t>(),
                                            v_out.as<uint32_t>(), static_cast<int>(n_post), 0, end_bit, st);
        if (e != cudaSuccess) return fail(cuda_fail(e, "bm25 doc-row sort", __FILE__, __LINE__));
        bm25_drow_off_kernel<<<static_cast<unsigned>((n_docs + 1 + 255) / 256), 256, 0, st>>>(k_out.as<uint32_t>(), n_post, n_docs, bm->drow_off);
        count_launch();
        uint32_t h_nsp = 0;
        e = cudaMemcpyAsync(&h_nsp, bm->drow_off + n_docs, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail(cuda_fail(e, "bm25 doc-row offsets", __FILE__, __LINE__));
        n_sp = h_nsp;
        if (cudaMalloc(&bm->drow_term, std::max<size_t>(n_sp, 1) * 4) != cudaSuccess ||
            cudaMalloc(&bm->drow_val, std::max<size_t>(n_sp, 1) * 8) != cudaSuccess) {
          cudaGetLastError(); set_error("arb_bm25_create: out of device memory (doc rows)"); return fail(ARB_ENOMEM);
        }
        if (n_sp > 0) {
          bm25_drow_fill_kernel<<<static_cast<unsigned>((n_sp + 255) / 256), 256, 0, st>>>(
              v_out.as<uint32_t>(), n_sp, d_ptr.as<int64_t>(), n_terms, d_code.as<int32_t>(), bm->idf, bm->gtab, bm->drow_term, bm->drow_val);
          count_launch();
        }
        // the tile-major, term-major-inside-a-tile copy (k_in still holds the unsorted doc keys in term-major order)
        {
          const int64_t n_tiles = bm->dev.n_tiles;
          if (cudaMalloc(&bm->tt_off, static_cast<size_t>(n_tiles + 1) * 4) != cudaSuccess ||
              cudaMalloc(&bm->tt_term, std::max<size_t>(n_sp, 1) * 4) != cudaSuccess ||
              cudaMalloc(&bm->tt_doc, std::max<size_t>(n_sp, 1) + 16) != cudaSuccess ||
              cudaMalloc(&bm->tt_info, std::max<size_t>(n_sp, 1) * 4) != cudaSuccess ||
              cudaMalloc(&bm->tt_val, std::max<size_t>(n_sp, 1) * 8) != cudaSuccess) {
            cudaGetLastError(); set_error("arb_bm25_create: out of device memory (tile rows)"); return fail(ARB_ENOMEM);
          }
          const size_t bloom_bytes = static_cast<size_t>(n_tiles) * PLANE_TILE * BLOOM_W * 4;
          if (cudaMalloc(&bm->tt_bloom, bloom_bytes) != cudaSuccess) {
            cudaGetLastError(); set_error("arb_bm25_create: out of device memory (doc filters)"); return fail(ARB_ENOMEM);
          }
          e = cudaMemsetAsync(bm->tt_bloom, 0, bloom_bytes, st);
          if (e != cudaSuccess) return fail(cuda_fail(e, "bm25 doc filters", __FILE__, __LINE__));
          AsyncBuf t_key;
          e = t_key.alloc(np4, st);
          if (e != cudaSuccess) { cudaGetLastError(); set_error("arb_bm25_create: out of device memory (tile-row build)"); return fail(ARB_ENOMEM); }
          bm25_tile_key_kernel<<<static_cast<unsigned>((n_post + 255) / 256), 256, 0, st>>>(
              k_in.as<uint32_t>(), n_post, static_cast<uint32_t>(n_docs), static_cast<uint32_t>(n_tiles), t_key.as<uint32_t>());
          count_launch();
          int tbits = 1;
          while ((int64_t(1) << tbits) <= n_tiles) ++tbits;     // keys are in [0, n_tiles]
          size_t tb2 = 0;
          e = cub::DeviceRadixSort::SortPairs(nullptr, tb2, t_key.as<uint32_t>(), k_out.as<uint32_t>(), v_in.as<uint32_t>(),
                                              v_out.as<uint32_t>(), static_cast<int>(n_post), 0, tbits, st);
          AsyncBuf tmp2;
          if (e == cudaSuccess) e = tmp2.alloc(tb2, st);
          if (e == cudaSuccess)
            e = cub::DeviceRadixSort::SortPairs(tmp2.p, tb2, t_key.as<uint32_t>(), k_out.as<uint32_t>(), v_in.as<uint32_t>(),
                                                v_out.as<uint32_t>(), static_cast<int>(n_post), 0, tbits, st);
          if (e != cudaSuccess) return fail(cuda_fail(e, "bm25 tile-row sort", __FILE__, __LINE__));
          bm25_drow_off_kernel<<<static_cast<unsigned>((n_tiles + 1 + 255) / 256), 256, 0, st>>>(k_out.as<uint32_t>(), n_post, n_tiles, bm->tt_off);
          count_launch();
          if (n_sp > 0) {
            bm25_tt_fill_kernel<<<static_cast<unsigned>((n_sp + 255) / 256), 256, 0, st>>>(
                v_out.as<uint32_t>(), n_sp, d_ptr.as<int64_t>(), n_terms, d_doc.as<int32_t>(), d_code.as<int32_t>(), bm->idf, bm->gtab,
                bm->tt_term, bm->tt_doc, bm->tt_val, bm->tt_bloom);
            count_launch();
            bm25_tt_info_kernel<<<static_cast<unsigned>(n_tiles), 256, 0, st>>>(bm->tt_off, bm->tt_term, bm->tt_doc, bm->tt_info);
            count_launch();
          }
        }
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail(cuda_fail(e, "bm25 doc rows", __FILE__, __LINE__));
      } else {
        e = cudaMemsetAsync(bm->drow_off, 0, static_cast<size_t>(n_docs + 1) * 4, st);
        if (e == cudaSuccess && cudaMalloc(&bm->tt_off, static_cast<size_t>(bm->dev.n_tiles + 1) * 4) != cudaSuccess) e = cudaErrorMemoryAllocation;
        if (e == cudaSuccess) e = cudaMemsetAsync(bm->tt_off, 0, static_cast<size_t>(bm->dev.n_tiles + 1) * 4, st);
        if (e == cudaSuccess && cudaMalloc(&bm->tt_bloom, static_cast<size_t>(bm->dev.n_tiles) * PLANE_TILE * BLOOM_W * 4) != cudaSuccess) e = cudaErrorMemoryAllocation;
        if (e == cudaSuccess) e = cudaMemsetAsync(bm->tt_bloom, 0, static_cast<size_t>(bm->dev.n_tiles) * PLANE_TILE * BLOOM_W * 4, st);
        if (e != cudaSuccess) return fail(cuda_fail(e, "bm25 doc rows", __FILE__, __LINE__));
      }
      bm->dev.n_sp = n_sp;
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(cuda_fail(e, "bm25 build kernels", __FILE__, __LINE__));
  }
  bm->df.resize(n_terms);
  for (int64_t t = 0; t < n_terms; ++t) bm->df[t] = term_ptr[t + 1] - term_ptr[t];
  bm->dev.n_docs = n_docs; bm->dev.n_terms = n_terms; bm->dev.n_post = n_post;
  bm->dev.nb = nb; bm->dev.n_g = static_cast<int>(n_g);
  bm->dev.blk_off = bm->blk_off; bm->dev.post = bm->post; bm->dev.idf = bm->idf; bm->dev.gtab = bm->gtab;
  bm->dev.dense_slot = bm->dense_slot; bm->dev.fine_off = bm->fine_off;
  bm->dev.plane_slot = bm->plane_slot; bm->dev.planes = bm->planes;
  bm->dev.plane_tmax = bm->plane_tmax;
  bm->dev.plane_gmax = bm->plane_tmax ? bm->plane_tmax + static_cast<size_t>(bm->dev.n_tiles) * bm->dev.n_planes : nullptr;
  bm->dev.planes32 = bm->planes32;
  bm->dev.drow_off = bm->drow_off; bm->dev.drow_term = bm->drow_term; bm->dev.drow_val = bm->drow_val;
  bm->dev.tt_off = bm->tt_off; bm->dev.tt_term = bm->tt_term; bm->dev.tt_doc = bm->tt_doc; bm->dev.tt_info = bm->tt_info; bm->dev.tt_val = bm->tt_val;
  bm->dev.tt_bloom = bm->tt_bloom;
  *out = bm;
  return ARB_OK;
}

extern "C" int arb_bm25_destroy(arb_bm25* bm) {
  if (!bm) return ARB_OK;
  if (bm->blk_off) cudaFree(bm->blk_off);
  if (bm->post) cudaFree(bm->post);
  if (bm->idf) cudaFree(bm->idf);
  if (bm->gtab) cudaFree(bm->gtab);
  if (bm->dense_slot) cudaFree(bm->dense_slot);
  if (bm->fine_off) cudaFree(bm->fine_off);
  if (bm->plane_slot) cudaFree(bm->plane_slot);
  if (bm->planes) cudaFree(bm->planes);
  if (bm->plane_tmax) cudaFree(bm->plane_tmax);
  if (bm->planes32) cudaFree(bm->planes32);
  if (bm->drow_off) cudaFree(bm->drow_off);
  if (bm->drow_term) cudaFree(bm->drow_term);
  if (bm->drow_val) cudaFree(bm->drow_val);
  if (bm->tt_off) cudaFree(bm->tt_off);
  if (bm->tt_term) cudaFree(bm->tt_term);
  if (bm->tt_doc) cudaFree(bm->tt_doc);
  if (bm->tt_info) cudaFree(bm->tt_info);
  if (bm->tt_val) cudaFree(bm->tt_val);
  if (bm->tt_bloom) cudaFree(bm->tt_bloom);
  delete bm;
  return ARB_OK;
}

extern "C" int64_t arb_bm25_ndocs(const arb_bm25* bm) { return bm ? bm->dev.n_docs : 0; }
extern "C" int64_t arb_bm25_npostings(const arb_bm25* bm) { return bm ? bm->dev.n_post : 0; }

extern "C" int64_t arb_bm25_query_postings(const arb_bm25* bm, const int32_t* q_terms, int64_t nq, int max_qt) {
  if (!bm || !q_terms) return 0;
  int64_t total = 0;
  for (int64_t i = 0; i < nq * max_qt; ++i) {
    const int32_t t = q_terms[i];
    if (t >= 0 && t < bm->dev.n_terms) total += bm->df[t];
  }
  return total;
}

static int stage_terms(const int32_t* q_terms, int64_t nq, int max_qt, int q_is_device, AsyncBuf& buf,
                       const int32_t** dev_terms, cudaStream_t st) {
  if (q_is_device) { *dev_terms = q_terms; return ARB_OK; }
  ARB_CUDA(buf.alloc(static_cast<size_t>(nq) * max_qt * 4, st));
  ARB_CUDA(cudaMemcpyAsync(buf.p, q_terms, static_cast<size_t>(nq) * max_qt * 4, cudaMemcpyHostToDevice, st));
  *dev_terms = buf.as<int32_t>();
  return ARB_OK;
}

extern "C" int arb_bm25_search(const arb_bm25* bm, const int32_t* q_terms, int64_t nq, int max_qt, int q_is_device,
                               int k, int64_t id_base, double* out_scores, int64_t* out_ids, int out_is_device,
                               void* stream) {
  ARB_CHECK(bm != nullptr, ARB_EINVAL, "arb_bm25_search: NULL index");
  ARB_CHECK(nq >= 0 && (q_terms != nullptr || nq == 0), ARB_EINVAL, "arb_bm25_search: bad q_terms/nq");
  ARB_CHECK(max_qt >= 1 && max_qt <= ARB_MAX_QUERY_TERMS, ARB_ELIMIT, "arb_bm25_search: max_qt must be in [1, 64]");
  ARB_CHECK(k >= 1 && k <= ARB_MAX_TOPK, ARB_ELIMIT, "arb_bm25_search: k must be in [1, 256]");
  ARB_CHECK(out_scores != nullptr && out_ids != nullptr, ARB_EINVAL, "arb_bm25_search: NULL output");
  ARB_CHECK(nq <= 2147483647ll, ARB_ELIMIT, "arb_bm25_search: too many queries");
  if (nq == 0) return ARB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = bm25_set_smem_attrs();
  if (rc) return rc;
  AsyncBuf tbuf, pscores, pids, oscores, oids;
  const int32_t* d_terms = nullptr;
  rc = stage_terms(q_terms, nq, max_qt, q_is_device, tbuf, &d_terms, st);
  if (rc) return rc;

  double* d_scores = out_scores;
  int64_t* d_ids = out_ids;
  if (!out_is_device) {
    ARB_CUDA(oscores.alloc(static_cast<size_t>(nq) * k * 8, st));
    ARB_CUDA(oids.alloc(static_cast<size_t>(nq) * k * 8, st));
    d_scores = oscores.as<double>();
    d_ids = oids.as<int64_t>();
  }

  // Batches go through the plane/sparse pipeline (bm25_batch.cu); small batches and long queries use the
  // per-query kernel.  ARB_BM25_MODE=query|batch forces one of them (tests exercise both).
  int64_t batch_min = 256;                                       // below this the tile-major pipeline's per-tile fixed costs lose to one CTA per query
  bool force_query = false;
  if (const char* v = std::getenv("ARB_BM25_MODE")) {
    if (!strcmp(v, "query")) force_query = true;
    else if (!strcmp(v, "batch")) batch_min = 1;
  }
  if (!force_query && nq >= batch_min && max_qt <= 32 && bm->dev.n_docs > 0) {
    ProfileScope pscope(PROF_BM25_TOPK, st);
    rc = bm25_batch_search(bm->dev, d_terms, nq, max_qt, k, id_base, d_scores, d_ids, st);
    if (rc) return rc;
  } else {
    // few queries: cut the doc blocks over several CTAs per query and merge
    const int sms = num_sms();
    int n_split = 1;
    if (nq < sms) n_split = static_cast<int>(std::min<int64_t>(bm->dev.nb, (sms + nq - 1) / nq));
    n_split = std::max(1, std::min(n_split, 8192 / k));
    const int bps = (bm->dev.nb + n_split - 1) / n_split;
    n_split = (bm->dev.nb + bps - 1) / bps;
    double* ps = d_scores;
    int64_t* pi = d_ids;
    if (n_split > 1) {
      ARB_CUDA(pscores.alloc(static_cast<size_t>(n_split) * nq * k * 8, st));
      ARB_CUDA(pids.alloc(static_cast<size_t>(n_split) * nq * k * 8, st));
      ps = pscores.as<double>();
      pi = pids.as<int64_t>();
    }
    {
      ProfileScope pscope(PROF_BM25_TOPK, st);
      rc = launch_bm25_topk(bm->dev, d_terms, max_qt, k, id_base, nq, n_split, bps, nullptr, nullptr,
                            static_cast<int>(nq), ps, pi, st);
      if (rc) return rc;
    }
    if (n_split > 1) {
      rc = merge_topk_device(ARB_F64, ps, pi, n_split, nq, k, d_scores, d_ids, st);
      if (rc) return rc;
    }
  }
  if (!out_is_device) {
    ARB_CUDA(cudaMemcpyAsync(out_scores, d_scores, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaMemcpyAsync(out_ids, d_ids, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaStreamSynchronize(st));
  }
  return ARB_OK;
}

extern "C" int arb_bm25_scores(const arb_bm25* bm, const int32_t* q_terms, int64_t nq, int max_qt, int q_is_device,
                               int normalise, int64_t n_out, double* out, int out_is_device, void* stream) {
  ARB_CHECK(bm != nullptr, ARB_EINVAL, "arb_bm25_scores: NULL index");
  ARB_CHECK(nq >= 0 && (q_terms != nullptr || nq == 0), ARB_EINVAL, "arb_bm25_scores: bad q_terms/nq");
  ARB_CHECK(max_qt >= 1 && max_qt <= ARB_MAX_QUERY_TERMS, ARB_ELIMIT, "arb_bm25_scores: max_qt must be in [1, 64]");
  ARB_CHECK(n_out >= 0 && (out != nullptr || n_out == 0), ARB_EINVAL, "arb_bm25_scores: bad output");
  ARB_CHECK(nq <= 65535 * 64ll, ARB_ELIMIT, "arb_bm25_scores: too many queries in one call");
  if (nq == 0 || n_out == 0) return ARB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = bm25_set_smem_attrs();
  if (rc) return rc;
  AsyncBuf tbuf, obuf;
  const int32_t* d_terms = nullptr;
  rc = stage_terms(q_terms, nq, max_qt, q_is_device, tbuf, &d_terms, st);
  if (rc) return rc;
  double* d_out = out;
  if (!out_is_device) { ARB_CUDA(obuf.alloc(static_cast<size_t>(nq) * n_out * 8, st)); d_out = obuf.as<double>(); }
  const int64_t total = nq * n_out;
  const int accumulate = (normalise >> 1) & 1;                 // bit 1: continue the sums already in `out` (queries of > 64 tokens)
  normalise &= 1;
  ARB_CHECK(!accumulate || out_is_device, ARB_EINVAL, "arb_bm25_scores: accumulation needs a device-resident output");
  if (n_out > bm->dev.n_docs && !accumulate) {                 // pad slots beyond the corpus with 0.0 (:320-323)
    fill_f64_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(d_out, 0.0, total);
    ARB_LAUNCH_OK();
  }
  const int nb_used = static_cast<int>(std::min<int64_t>(bm->dev.nb, (std::min(n_out, bm->dev.n_docs) + DOC_BLOCK - 1) / DOC_BLOCK));
  if (nb_used > 0) {
    dim3 grid(static_cast<unsigned>(nq), static_cast<unsigned>(nb_used));
    bm25_scores_kernel<<<grid, BM25_THREADS, BM25_SMEM, st>>>(bm->dev, d_terms, max_qt, n_out, d_out, accumulate);
    ARB_LAUNCH_OK();
  }
  if (normalise) {
    bm25_normalise_kernel<<<static_cast<unsigned>(nq), 1024, 0, st>>>(d_out, n_out);
    ARB_LAUNCH_OK();
  }
  if (!out_is_device) {
    ARB_CUDA(cudaMemcpyAsync(out, d_out, static_cast<size_t>(total) * 8, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaStreamSynchronize(st));
  }
  return ARB_OK;
}

extern "C" int arb_combine_fields(int n_fields, const double* const* field_scores, const double* weights, int64_t n,
                                  int64_t nq, double* out, void* stream) {
  ARB_CHECK(n_fields >= 1 && n_fields <= 8, ARB_ELIMIT, "arb_combine_fields: 1..8 fields");
  ARB_CHECK(field_scores && weights && out && n >= 0, ARB_EINVAL, "arb_combine_fields: bad argument");
  ARB_CHECK(nq == 0 || (nq > 0 && n % nq == 0), ARB_EINVAL, "arb_combine_fields: n must be a multiple of nq");
  if (n == 0) return ARB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FieldPtrs f;
  for (int k = 0; k < 8; ++k) { f.s[k] = k < n_fields ? field_scores[k] : nullptr; f.w[k] = k < n_fields ? weights[k] : 0.0; }
  combine_fields_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(f, n_fields, n, out);
  ARB_LAUNCH_OK();
  if (nq > 0) {
    bm25_normalise_kernel<<<static_cast<unsigned>(nq), 1024, 0, st>>>(out, n / nq);
    ARB_LAUNCH_OK();
  }
  return ARB_OK;
}
