// Batched BM25 top-k (sm_100a): the plane / sparse pipeline behind arb_bm25_search.
//
// Replaces corpus.get_scores() + top-k for a whole query batch (reference
// utils/bm25_search.py:309 -> rank_bm25 BM25Okapi.get_scores / SimpleBM25.get_scores :43-63).
//
// Why two passes.  In a Zipf corpus a few dozen terms own almost all postings (cfg3: the 64
// densest terms carry 98 % of the postings a query touches).  Scatter-adding those postings
// into shared-memory accumulators costs three 64-bit shared accesses each; instead every such
// term is stored as a doc-major fp64 "plane" of its finished contributions idf*g (+0.0 where
// the term is absent), and a CTA that has staged one 256-doc tile of all planes in shared
// memory scores that tile for EVERY query of the batch with register accumulators:
// one conflict-free LDS + one DADD per (query token, doc).  The fp64 addition order of the
// reference (query-token order per doc) is kept exactly:
//
//   1. sparse pass   (one warp per query)  docs that match at least one non-plane token
//      ("members") get their complete score in query-token order: postings of the sparse tokens
//      are scattered into a per-warp 1024-doc accumulator tile, plane tokens are looked up for the
//      members only.  The members are recorded in a per-query bitmap; their top-k (list A) gives
//      a lower bound on the final k-th score.
//   2. plane pass    (one CTA per 256-doc tile, all queries)  every other doc has plane tokens
//      only, so its score is the in-order sum of plane values; docs that reach A's k-th score
//      and are not members are appended to the query's list E.
//   3. final merge   A + E -> top-k by (score desc, doc asc).
//   4. fallback      queries whose A holds fewer than k docs while they have plane tokens (no
//      usable bound), or whose E overflowed, are re-scored by the per-query kernel (bm25.cu).
//
// Adding +0.0 for absent terms is what the reference itself does (score += idf * 0), and an
// accumulator that starts at +0.0 can never become -0.0, so the result is bit-identical.
#include "bm25.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace arb {

constexpr int SP_MIN_BLOCK_LOG2 = 10;                // sparse-pass doc blocks: 1024 .. 16384 docs, chosen per query
constexpr int SP_MAX_BLOCK_LOG2 = 14;
constexpr int SP_MAXB = 1 << SP_MAX_BLOCK_LOG2;
constexpr int SP_MAXM = 1024;                       // members per block visit (rank-compact accumulators)
constexpr int SP_TARGET_M = 448;                    // block size aims at this many members per visit
constexpr int SP_POOL = 1728;                       // fp64 pool: M accumulators + staged plane values
constexpr int SP_PCAP = 1024;                       // postings staged per visit
constexpr int SP_STAGED_PLANES = 4;                 // plane tokens whose member values are staged by cp.async
constexpr int PL_WARPS = 16;
constexpr int PL_SMEM_MAX = 227 * 1024 - 64;        // planes tile + per-query records (64 B left for the static mbarrier)
constexpr int E_CAP = 1024;                         // plane-pass emissions kept per query
constexpr unsigned long long THR_SKIP = ~0ull;
constexpr int H3_MAX = 1024;                        // queries per batch chunk that may take the histogram path
constexpr int H3_BINS = 4096;                       // 128 bins per octave from 2^-12 up
constexpr unsigned long long H3_TAG = 0xFFF8000000000000ull;   // qrec.x of a histogram query (a NaN pattern) | slot
constexpr unsigned long long H3_BASE = (1ull << 18) | (static_cast<unsigned long long>(1023 - 12) << 7);
// bin of a positive score: sign | exponent | 7 mantissa bits of its order-preserving key, clamped
__device__ __forceinline__ int h3_bin(double v) {
  const long long t = static_cast<long long>(f64_key(v) >> 45) - static_cast<long long>(H3_BASE);
  return static_cast<int>(t < 0 ? 0 : (t > H3_BINS - 1 ? H3_BINS - 1 : t));
}
// smallest key of bin b (b >= 1); bin 0 also holds everything below 2^-12, so its edge is "any positive score"
__device__ __forceinline__ unsigned long long h3_edge(int b) {
  return b == 0 ? f64_key(0.0) + 1ull : (static_cast<unsigned long long>(b) + H3_BASE) << 45;
}
static_assert(SP_POOL >= SP_MAXM, "the pool holds at least the accumulators");

// per-warp state of the sparse pass; CAP = candidate buffer (>= 2k + 32)
template <int CAP>
struct SpWarp {
  double pool[SP_POOL];                  // [0, M): accumulator of member r (ranked by doc id inside the block);
                                         // [M, ...): plane values of the members, staged by cp.async
  unsigned long long ckey[CAP];
  uint32_t cdoc[CAP];
  uint32_t post[SP_PCAP];                // this visit's postings of all sparse tokens, staged by cp.async
  uint32_t bm[SP_MAXB / 32];             // member bitmap of the block being visited
  uint16_t prk[SP_PCAP];                 // member rank of each staged posting (0xFFFF: other sub-block)
  uint16_t pref[SP_MAXB / 32];           // rank of the first member of each bitmap word
  uint16_t mdoc[SP_MAXM];                // doc offset (inside the block) of member r
};
template <int CAP> constexpr int sp_warps() { return CAP <= 256 ? 8 : 7; }

__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

struct BatchScratch {
  unsigned long long* a_key;   // [nq, k]  sparse-pass top-k (list A), sorted
  uint32_t* a_doc;             // [nq, k]
  int* a_cnt;                  // [nq]
  unsigned long long* thr;     // [nq]  key of A's k-th score, or THR_SKIP
  uint8_t* state;              // [nq]  0 = plane pass runs, 1 = no plane tokens (A is final), 2 = fallback
  uint8_t* qd;                 // [nq, qd_stride] the query's plane slots in token order, 0xFF terminated/padded
  int qd_stride;               // mq rounded up to a multiple of 8 (<= 32)
  uint32_t* bitmap;            // [bm_tiles, nq, 8] member bitmap, tile-major: the plane pass streams one tile's
                               // 32-byte records of consecutive queries (NULL when the index has no planes)
  int64_t bm_tiles;            // 256-doc tiles covered (whole 16K-doc blocks)
  int64_t nq;
  unsigned long long* e_key;   // [nq, E_CAP]
  uint32_t* e_doc;             // [nq, E_CAP]
  int* e_cnt;                  // [nq]
  int* flag_list;              // [nq]
  // the sparse pass cuts every query into `parts` doc ranges (whole 16K-doc blocks) scored by different
  // warps, so that one heavy query cannot become the tail of the kernel; part lists are merged into A
  int parts;
  unsigned long long* pa_key;  // [nq, parts, k]
  uint32_t* pa_doc;            // [nq, parts, k]
  int* pa_cnt;                 // [nq, parts]  entries, or -1 if the part overflowed its staging areas
  unsigned long long* gthr;    // [nq] largest k-th key any part of the query has reached so far (0 = none): a doc
                               // strictly below it cannot be in the query's top-k, whichever part it lies in
  int* counters;               // [0] next unit of the sparse pass, [1] number of flagged queries, [2] histogram queries
  // queries whose list A holds fewer than k docs while they have plane tokens get their plane-pass threshold from
  // a histogram of the plane-only scores of all their non-member docs (bm25_plane_list_kernel)
  int* h3_list;                // [H3_MAX]
  uint32_t* hist;              // [H3_MAX, H3_BINS]
};

// bitonic sort of n_pow2 (key, doc) pairs by (key desc, doc asc); one warp
__device__ __forceinline__ void warp_sort_desc(unsigned long long* key, uint32_t* doc, int n_pow2, int lane) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncwarp();
      for (int i = lane; i < n_pow2 / 2; i += 32) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool best_first = ((lo & size) == 0);
        const unsigned long long ka = key[lo], kb = key[hi];
        const uint32_t da = doc[lo], db = doc[hi];
        const bool a_before_b = (ka > kb) || (ka == kb && da < db);
        if (a_before_b != best_first) { key[lo] = kb; key[hi] = ka; doc[lo] = db; doc[hi] = da; }
      }
    }
  }
  __syncwarp();
}

// same, whole CTA
template <class IdT>
__device__ __forceinline__ void cta_sort_desc(unsigned long long* key, IdT* id, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n_pow2 / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool best_first = ((lo & size) == 0);
        const unsigned long long ka = key[lo], kb = key[hi];
        const IdT da = id[lo], db = id[hi];
        const bool a_before_b = (ka > kb) || (ka == kb && da < db);
        if (a_before_b != best_first) { key[lo] = kb; key[hi] = ka; id[lo] = db; id[hi] = da; }
      }
    }
  }
  __syncthreads();
}

// ----------------------------------------------------------------------------
// 1. sparse pass: one warp per query (queries handed out by an atomic counter).
//
// The warp walks the corpus in doc blocks of B = 1024 .. 8192 docs (B chosen per query so that a
// block holds ~SP_TARGET_M members on average).  The pass is bound by memory latency (a few hundred
// scattered postings and plane values per visit), so every global read of a visit is issued as one
// burst of cp.async copies into the warp's shared-memory staging areas:
//  (1) the postings of all the query's sparse tokens in this block are staged, then mark the member
//      bitmap;
//  (2) members are ranked (bitmap word prefix sums): the fp64 accumulators are indexed by rank, so
//      their footprint does not depend on B; every staged posting records its member's rank;
//  (3) the plane values of the members are staged in one burst (member-major: one address
//      computation per member);
//  (4) tokens are applied in QUERY ORDER: a sparse token adds idf*g at the rank of each of its docs,
//      a plane token adds its staged value for every member;
//  (5) members are harvested in ascending doc order.
// ----------------------------------------------------------------------------
template <int CAP>
__global__ void __launch_bounds__(sp_warps<CAP>() * 32, 1)
bm25_sparse_pass_kernel(const Bm25Dev ix, const int32_t* __restrict__ q_terms_all, int mq, int64_t nq, int k,
                        const BatchScratch sc) {
  extern __shared__ __align__(16) uint8_t sp_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  SpWarp<CAP>& W = reinterpret_cast<SpWarp<CAP>*>(sp_smem)[warp];
  for (int i = lane; i < SP_MAXB / 32; i += 32) W.bm[i] = 0u;    // restored to zero after every block
  __syncwarp();
  const int nsl = ix.n_slices;
  const uint32_t lt = (1u << lane) - 1u;
  const unsigned long long key_zero = f64_key(0.0);
  const int64_t tile_stride = static_cast<int64_t>(ix.n_planes) * PLANE_TILE;

  while (true) {
    int ui = 0;
    if (lane == 0) ui = atomicAdd(&sc.counters[0], 1);
    const int64_t unit = __shfl_sync(0xffffffffu, ui, 0);
    if (unit >= nq * sc.parts) break;
    // part-major order: the warps in flight work on the same doc range of different queries, so the term planes
    // of that range (n_planes * 8 B per doc) stay in L2 for the scattered member look-ups
    const int part = static_cast<int>(unit / nq);
    const int64_t q = unit - static_cast<int64_t>(part) * nq;

    // one query token per lane: kind 1 = plane term, 2 = sparse term with per-512-doc offsets,
    // 3 = sparse term with per-16K-block offsets only, 0 = contributes nothing
    int t = -1, kind = 0, slot = -1;
    double idf = 0.0;
    long long df = 0;
    if (lane < mq) {
      t = q_terms_all[q * mq + lane];
      if (t >= 0 && t < ix.n_terms) {
        idf = __ldg(ix.idf + t);
        if (idf != 0.0) {                                      // (idf.get(q) or 0) == 0 adds +0.0: a no-op
          const int ps = ix.n_planes > 0 ? __ldg(ix.plane_slot + t) : -1;
          if (ps >= 0) { kind = 1; slot = ps; }
          else {
            const int fs = __ldg(ix.dense_slot + t);
            if (fs >= 0) { kind = 2; slot = fs; } else kind = 3;
            df = static_cast<long long>(__ldg(ix.blk_off + static_cast<int64_t>(t + 1) * ix.nb)) -
                 static_cast<long long>(__ldg(ix.blk_off + static_cast<int64_t>(t) * ix.nb));
          }
        }
      }
    }
    const uint32_t plane_mask = __ballot_sync(0xffffffffu, kind == 1);
    const uint32_t sparse_mask = __ballot_sync(0xffffffffu, kind >= 2);
    const int np = __popc(plane_mask);
    // the plane pass reads the query's plane slots in token order, packed at the front of an 8-byte-aligned row
    if (part == 0) {
      if (kind == 1) sc.qd[q * sc.qd_stride + __popc(plane_mask & lt)] = static_cast<uint8_t>(slot);
      if (lane >= np && lane < sc.qd_stride) sc.qd[q * sc.qd_stride + lane] = 0xFF;
    }
    // plane offsets (slot * 256) of the first plane tokens, for the member-major staging loop
    int64_t pso[SP_STAGED_PLANES];
#pragma unroll
    for (int p = 0; p < SP_STAGED_PLANES; ++p) {
      const int jl = p < np ? static_cast<int>(__fns(plane_mask, 0, p + 1)) : 0;
      pso[p] = static_cast<int64_t>(__shfl_sync(0xffffffffu, slot, jl)) * PLANE_TILE;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) df += __shfl_xor_sync(0xffffffffu, df, o);
    // block size: the largest power of two in [1024, 8192] with (expected members per block) <= SP_TARGET_M
    int bl2 = SP_MAX_BLOCK_LOG2;
    while (bl2 > SP_MIN_BLOCK_LOG2 && (df << bl2) > static_cast<long long>(SP_TARGET_M) * ix.n_docs) --bl2;
    const int nblk_all = static_cast<int>((ix.n_docs + (int64_t(1) << bl2) - 1) >> bl2);
    // this part's blocks: 16K-doc blocks [part * nb / parts, (part + 1) * nb / parts)
    const int b_first = static_cast<int>((static_cast<int64_t>(part) * ix.nb / sc.parts) << (DOC_BLOCK_LOG2 - bl2));
    const int nblk = min(nblk_all, static_cast<int>((static_cast<int64_t>(part + 1) * ix.nb / sc.parts) << (DOC_BLOCK_LOG2 - bl2)));
    const int nwords = 1 << (bl2 - 5);
    const int wpl = nwords >> 5;                               // bitmap words ranked by each lane (contiguous)
    const uint32_t sub_mask = (1u << (DOC_BLOCK_LOG2 - bl2)) - 1u;
    const uint32_t in_mask = (1u << bl2) - 1u;
    const int spb = 1 << (bl2 - SLICE_LOG2);                   // 512-doc slices per block

    unsigned long long thr_key = key_zero;                     // only positive scores are candidates
    unsigned long long floor_key = *reinterpret_cast<volatile unsigned long long*>(sc.gthr + q);   // other parts' bound
    int cnt = 0;
    bool overflow = false;
    // word w of this query's member bitmap lives at bitmap[((w >> 3) * nq + q) * 8 + (w & 7)]
    uint32_t* bm_row = sc.bitmap ? sc.bitmap + q * 8 : nullptr;
    const int64_t bm_tile_stride = sc.nq * 8;

    auto compact = [&]() {
      __syncwarp();
      int n2 = 32;
      while (n2 < cnt) n2 <<= 1;
      for (int i = cnt + lane; i < n2; i += 32) { W.ckey[i] = 0ull; W.cdoc[i] = 0xFFFFFFFFu; }
      warp_sort_desc(W.ckey, W.cdoc, n2, lane);
      if (cnt >= k) {
        cnt = k; thr_key = W.ckey[k - 1];
        if (lane == 0 && thr_key > floor_key) atomicMax(sc.gthr + q, thr_key);
      }
      __syncwarp();
    };
    // posting range of this lane's token in block b (kind 3: the enclosing 16K block, fetched once per 16K docs)
    auto fetch_range = [&](int b, uint32_t& lo, uint32_t& hi) {
      if (b >= nblk) return;
      if (kind == 2) {
        const uint32_t* fo = ix.fine_off + static_cast<int64_t>(slot) * (nsl + 1);
        lo = __ldg(fo + min(b * spb, nsl));
        hi = __ldg(fo + min(b * spb + spb, nsl));
      } else if (kind == 3 && (static_cast<uint32_t>(b) & sub_mask) == 0u) {
        const int64_t o = static_cast<int64_t>(t) * ix.nb + (b >> (DOC_BLOCK_LOG2 - bl2));
        lo = __ldg(ix.blk_off + o);
        hi = __ldg(ix.blk_off + o + 1);
      }
    };

    uint32_t lo = 0u, hi = 0u, nlo = 0u, nhi = 0u;
    if (sparse_mask) fetch_range(b_first, nlo, nhi);
    for (int b = b_first; b < nblk && sparse_mask; ++b) {
      lo = nlo; hi = nhi;
      fetch_range(b + 1, nlo, nhi);                            // in flight during this visit
      const uint32_t sub = static_cast<uint32_t>(b) & sub_mask;
      const uint32_t ne = __ballot_sync(0xffffffffu, hi > lo) & sparse_mask;
      if (ne == 0u) {
        if (bm_row)
          for (int i = lane; i < nwords; i += 32) {
            const int64_t w = static_cast<int64_t>(b) * nwords + i;
            bm_row[(w >> 3) * bm_tile_stride + (w & 7)] = 0u;
          }
        continue;
      }

      floor_key = max(floor_key, *reinterpret_cast<volatile unsigned long long*>(sc.gthr + q));

      // ---- (1) the active tokens are packed into lanes 0..na-1; their postings are staged back to back
      const int na = __popc(ne);
      const int src = lane < na ? static_cast<int>(__fns(ne, 0, lane + 1)) : 0;
      const uint32_t c_lo = __shfl_sync(0xffffffffu, lo, src);
      const uint32_t c_hi = __shfl_sync(0xffffffffu, hi, src);
      const uint32_t c_len = lane < na ? c_hi - c_lo : 0u;
      uint32_t c_end = c_len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, c_end, o);
        if (lane >= o) c_end += v;
      }
      const uint32_t c_st = c_end - c_len;
      const uint32_t L = __shfl_sync(0xffffffffu, c_end, 31);
      if (L > SP_PCAP) { overflow = true; break; }
      for (int a = 0; a < na; ++a) {
        const uint32_t l = __shfl_sync(0xffffffffu, c_lo, a), n = __shfl_sync(0xffffffffu, c_len, a);
        const uint32_t st = __shfl_sync(0xffffffffu, c_st, a);
#pragma unroll 2
        for (uint32_t i = lane; i < n; i += 32) cp_async_4(&W.post[st + i], ix.post + l + i);
      }
      cp_async_wait_all();
      __syncwarp();
      for (uint32_t f = lane; f < L; f += 32) {
        const uint32_t d16 = post_doc(W.post[f]);
        if ((d16 >> bl2) == sub) atomicOr(&W.bm[(d16 & in_mask) >> 5], 1u << (d16 & 31));
      }
      __syncwarp();

      // ---- (2) rank the members
      int c = 0;
      for (int i = 0; i < wpl; ++i) c += __popc(W.bm[lane * wpl + i]);
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int M = __shfl_sync(0xffffffffu, incl, 31);
      if (bm_row)
        for (int i = lane; i < nwords; i += 32) {
          const int64_t w = static_cast<int64_t>(b) * nwords + i;
          bm_row[(w >> 3) * bm_tile_stride + (w & 7)] = W.bm[i];
        }
      if (M == 0) continue;                                    // the 16K block's postings lie in other sub-blocks
      if (M > SP_MAXM) {                                       // denser than the staging areas allow: fallback
        overflow = true;
        for (int i = lane; i < nwords; i += 32) W.bm[i] = 0u;
        __syncwarp();
        break;
      }
      {
        int r = incl - c;
        for (int i = 0; i < wpl; ++i) {
          const int wi = lane * wpl + i;
          W.pref[wi] = static_cast<uint16_t>(r);
          r += __popc(W.bm[wi]);
        }
      }
      __syncwarp();
      for (uint32_t f = lane; f < L; f += 32) {
        const uint32_t d16 = post_doc(W.post[f]);
        uint32_t rk = 0xFFFFu;
        if ((d16 >> bl2) == sub) {
          const uint32_t d = d16 & in_mask;
          rk = W.pref[d >> 5] + __popc(W.bm[d >> 5] & ((1u << (d & 31)) - 1u));
          W.mdoc[rk] = static_cast<uint16_t>(d);               // every posting of the doc writes the same value
        }
        W.prk[f] = static_cast<uint16_t>(rk);
      }
      for (int i = lane; i < M; i += 32) W.pool[i] = 0.0;
      __syncwarp();
      const int64_t doc0 = static_cast<int64_t>(b) << bl2;

      // ---- (3) stage the plane values of the members (first SP_STAGED_PLANES plane tokens, pool permitting)
      const int n_staged = min(min(np, SP_STAGED_PLANES), (SP_POOL - M) / M);
      if (n_staged > 0) {
#pragma unroll 2
        for (int i = lane; i < M; i += 32) {
          const int64_t doc = doc0 + W.mdoc[i];
          const double* src_p = ix.planes + (doc >> PLANE_TILE_LOG2) * tile_stride + plane_pos(static_cast<uint32_t>(doc) & (PLANE_TILE - 1));
          double* dst = W.pool + M + i;
#pragma unroll
          for (int p = 0; p < SP_STAGED_PLANES; ++p)
            if (p < n_staged) cp_async_8(dst + p * M, src_p + pso[p]);
        }
        cp_async_wait_all();
        __syncwarp();
      }

      // ---- (4) accumulate in query-token order
      int p_seen = 0;
      for (uint32_t m = ne | plane_mask; m; m &= m - 1) {
        const int jl = __ffs(m) - 1;
        if ((plane_mask >> jl) & 1u) {
          if (p_seen < n_staged) {
            const double* stg = W.pool + static_cast<size_t>(p_seen + 1) * M;
            for (int i = lane; i < M; i += 32) W.pool[i] = __dadd_rn(W.pool[i], stg[i]);
          } else {
            const double* pp = ix.planes + static_cast<int64_t>(__shfl_sync(0xffffffffu, slot, jl)) * PLANE_TILE;
            for (int i0 = 0; i0 < M; i0 += 128) {
              double v[4];
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const int i = i0 + r * 32 + lane;
                v[r] = 0.0;
                if (i < M) {
                  const int64_t doc = doc0 + W.mdoc[i];
                  v[r] = __ldg(pp + (doc >> PLANE_TILE_LOG2) * tile_stride + plane_pos(static_cast<uint32_t>(doc) & (PLANE_TILE - 1)));
                }
              }
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const int i = i0 + r * 32 + lane;
                if (i < M) W.pool[i] = __dadd_rn(W.pool[i], v[r]);
              }
            }
          }
          ++p_seen;
        } else {
          const int a = __popc(ne & ((1u << jl) - 1u));         // index of this token among the active ones
          const uint32_t st = __shfl_sync(0xffffffffu, c_st, a), en = __shfl_sync(0xffffffffu, c_end, a);
          const double w_idf = __shfl_sync(0xffffffffu, idf, jl);
          for (uint32_t f = st + lane; f < en; f += 32) {
            const uint32_t rk = W.prk[f];
            if (rk != 0xFFFFu)
              W.pool[rk] = __dadd_rn(W.pool[rk], __dmul_rn(w_idf, __ldg(ix.gtab + post_code(W.post[f]))));
          }
        }
        __syncwarp();
      }

      // ---- (5) harvest the members (ascending doc id, so a tie with the k-th score loses: strict >)
      for (int i0 = 0; i0 < M; i0 += 32) {
        const int i = i0 + lane;
        const bool valid = i < M;
        double v = 0.0;
        uint32_t d = 0u;
        if (valid) { v = W.pool[i]; d = W.mdoc[i]; }
        const unsigned long long key = f64_key(v);
        const bool pass = valid && key > thr_key && key >= floor_key;
        const uint32_t pm = __ballot_sync(0xffffffffu, pass);
        if (pass) {
          const int pos = cnt + __popc(pm & lt);
          W.ckey[pos] = key;
          W.cdoc[pos] = static_cast<uint32_t>(doc0) + d;
        }
        cnt += __popc(pm);
        if (cnt > CAP - 32) compact();
      }
      for (int i = lane; i < nwords; i += 32) W.bm[i] = 0u;
      __syncwarp();
    }
    if (bm_row && !sparse_mask)
      for (int64_t w = static_cast<int64_t>(b_first) * nwords + lane; w < static_cast<int64_t>(nblk) * nwords; w += 32)
        bm_row[(w >> 3) * bm_tile_stride + (w & 7)] = 0u;

    compact();
    cnt = min(cnt, k);
    const int64_t po = (q * sc.parts + part) * k;
    for (int i = lane; i < cnt; i += 32) { sc.pa_key[po + i] = W.ckey[i]; sc.pa_doc[po + i] = W.cdoc[i]; }
    if (lane == 0) sc.pa_cnt[q * sc.parts + part] = overflow ? -1 : cnt;
    __syncwarp();
  }
}

// ----------------------------------------------------------------------------
// 1b. merge the part lists of every query into list A (sorted top-k of the members), and derive the
// plane-pass threshold / the query's state.  One CTA per query.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
bm25_a_merge_kernel(int64_t nq, int k, const int A_MERGE_CAP, const BatchScratch sc) {
  extern __shared__ __align__(16) uint8_t am_smem[];             // A_MERGE_CAP x 12 B (a power of two >= parts * k)
  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x;
  __shared__ int s_n, s_ovf, s_total;
  __shared__ unsigned long long s_floor;
  if (tid == 0) {
    // Every part list is sorted; a full part's k-th key is a lower bound of the merged k-th key, so only
    // entries at or above the largest such bound can be in the merged top-k.
    int ovf = 0, total = 0;
    unsigned long long fl = 0ull;
    for (int p = 0; p < sc.parts; ++p) {
      const int c = sc.pa_cnt[q * sc.parts + p];
      if (c < 0) { ovf = 1; continue; }
      total += c;
      if (c >= k) fl = max(fl, sc.pa_key[(q * sc.parts + p) * k + k - 1]);
    }
    s_n = 0; s_ovf = ovf; s_floor = fl; s_total = total;
  }
  __syncthreads();
  const unsigned long long fl = s_floor;
  unsigned long long* key = reinterpret_cast<unsigned long long*>(am_smem);
  uint32_t* doc = reinterpret_cast<uint32_t*>(am_smem + static_cast<size_t>(A_MERGE_CAP) * 8);
  for (int i = tid; i < sc.parts * k; i += blockDim.x) {
    const int p = i / k, j = i - p * k;
    if (j < sc.pa_cnt[q * sc.parts + p]) {
      const unsigned long long kk = sc.pa_key[(q * sc.parts + p) * k + j];
      if (kk >= fl) {
        const int pos = atomicAdd(&s_n, 1);
        if (pos < A_MERGE_CAP) { key[pos] = kk; doc[pos] = sc.pa_doc[(q * sc.parts + p) * k + j]; }
      }
    }
  }
  __syncthreads();
  if (s_n > A_MERGE_CAP) { if (tid == 0) s_ovf = 1; }
  const int n = min(s_n, A_MERGE_CAP);
  __syncthreads();
  int n2 = 2;
  while (n2 < n) n2 <<= 1;
  for (int i = n + tid; i < n2; i += blockDim.x) { key[i] = 0ull; doc[i] = 0xFFFFFFFFu; }
  cta_sort_desc(key, doc, n2);
  const int cnt = min(n, k);
  for (int i = tid; i < k; i += blockDim.x) {
    sc.a_key[q * k + i] = i < cnt ? key[i] : 0ull;
    sc.a_doc[q * k + i] = i < cnt ? doc[i] : 0xFFFFFFFFu;
  }
  if (tid == 0) {
    const bool has_planes = sc.qd[q * sc.qd_stride] != 0xFF;
    sc.a_cnt[q] = cnt;
    int h3 = H3_MAX;
    if (!has_planes && !s_ovf)   { sc.state[q] = 1; sc.thr[q] = THR_SKIP; }
    else if (cnt >= k && !s_ovf) { sc.state[q] = 0; sc.thr[q] = key[k - 1]; }
    else if (!s_ovf && (h3 = atomicAdd(&sc.counters[2], 1)) < H3_MAX) {
      sc.state[q] = 3; sc.thr[q] = THR_SKIP; sc.h3_list[h3] = static_cast<int>(q);
    } else {
      sc.state[q] = 2; sc.thr[q] = THR_SKIP;
      sc.flag_list[atomicAdd(&sc.counters[1], 1)] = static_cast<int>(q);
    }
  }
}

// 1b'. the same merge for the common shapes (parts * k keys fit 20 KB): one WARP per query.  The part lists are
// sorted, so the merged k-th key is found by a bitwise search in which lane p counts, by binary search in its
// own list (held in shared memory), how many of part p's keys reach the candidate; only the <= k (+ ties)
// entries at or above that key are gathered and sorted.
constexpr int AMW_WARPS = 4;
constexpr int AMW_KEYS = 2560;                      // keys per warp (20 KB)
constexpr int AMW_OUT = 512;                        // gathered entries per warp
struct AmWarp { unsigned long long keys[AMW_KEYS]; unsigned long long okey[AMW_OUT]; uint32_t odoc[AMW_OUT]; };
__global__ void __launch_bounds__(AMW_WARPS * 32)
bm25_a_merge_warp_kernel(int64_t nq, int k, const BatchScratch sc) {
  extern __shared__ __align__(16) uint8_t amw_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q = blockIdx.x * static_cast<int64_t>(AMW_WARPS) + warp;
  if (q >= nq) return;
  AmWarp& W = reinterpret_cast<AmWarp*>(amw_smem)[warp];
  const int parts = sc.parts;
  int my_cnt = 0;
  bool ovf = false;
  if (lane < parts) { my_cnt = sc.pa_cnt[q * parts + lane]; if (my_cnt < 0) { ovf = true; my_cnt = 0; } }
  ovf = __any_sync(0xffffffffu, ovf);
  for (int p = 0; p < parts; ++p) {
    const int c = __shfl_sync(0xffffffffu, my_cnt, p);
    for (int j = lane; j < c; j += 32) W.keys[p * k + j] = sc.pa_key[(q * parts + p) * k + j];
  }
  __syncwarp();
  const int total = __reduce_add_sync(0xffffffffu, my_cnt);
  // number of keys of this lane's part that are >= cand (the list is sorted descending)
  auto count_ge = [&](unsigned long long cand) {
    int lo = 0, hi = my_cnt;
    const unsigned long long* L = W.keys + lane * k;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (L[mid] >= cand) lo = mid + 1; else hi = mid; }
    return lo;
  };
  unsigned long long T = 0ull;
  if (total > k) {
    for (int bit = 63; bit >= 0; --bit) {
      const unsigned long long cand = T | (1ull << bit);
      const int c = __reduce_add_sync(0xffffffffu, lane < parts ? count_ge(cand) : 0);
      if (c >= k) T = cand;                                    // at least k keys reach cand: the k-th key does too
    }
  }
  // gather everything at or above the k-th key (exactly k entries unless keys tie there)
  const int mine = lane < parts ? (total > k ? count_ge(T) : my_cnt) : 0;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  const int n = __shfl_sync(0xffffffffu, incl, 31);
  const bool too_many = n > AMW_OUT;
  if (!too_many) {
    for (int p = 0; p < parts; ++p) {
      const int c = __shfl_sync(0xffffffffu, mine, p), base = __shfl_sync(0xffffffffu, incl - mine, p);
      for (int j = lane; j < c; j += 32) { W.okey[base + j] = W.keys[p * k + j]; W.odoc[base + j] = sc.pa_doc[(q * parts + p) * k + j]; }
    }
  }
  int n2 = 32;
  while (n2 < n && n2 < AMW_OUT) n2 <<= 1;
  const int nn = too_many ? 0 : n;
  for (int i = nn + lane; i < n2; i += 32) { W.okey[i] = 0ull; W.odoc[i] = 0xFFFFFFFFu; }
  warp_sort_desc(W.okey, W.odoc, n2, lane);
  const int cnt = min(nn, k);
  for (int i = lane; i < k; i += 32) {
    sc.a_key[q * k + i] = i < cnt ? W.okey[i] : 0ull;
    sc.a_doc[q * k + i] = i < cnt ? W.odoc[i] : 0xFFFFFFFFu;
  }
  if (lane == 0) {
    const bool has_planes = sc.qd[q * sc.qd_stride] != 0xFF;
    const bool bad = ovf || too_many;
    sc.a_cnt[q] = cnt;
    int h3 = H3_MAX;
    if (!has_planes && !bad)   { sc.state[q] = 1; sc.thr[q] = THR_SKIP; }
    else if (cnt >= k && !bad) { sc.state[q] = 0; sc.thr[q] = W.okey[k - 1]; }
    else if (!bad && (h3 = atomicAdd(&sc.counters[2], 1)) < H3_MAX) {
      sc.state[q] = 3; sc.thr[q] = THR_SKIP; sc.h3_list[h3] = static_cast<int>(q);
    } else {
      sc.state[q] = 2; sc.thr[q] = THR_SKIP;
      sc.flag_list[atomicAdd(&sc.counters[1], 1)] = static_cast<int>(q);
    }
  }
}

// ----------------------------------------------------------------------------
// 2. plane pass: grid (tile sets, query groups).  The CTA stages one 256-doc tile of every
// plane in shared memory; each warp then scores that tile for its queries with 8 register
// accumulators per lane (lane l owns docs 8l .. 8l+7; see plane_pos()).
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(PL_WARPS * 32, 1)
bm25_plane_pass_kernel(const Bm25Dev ix, int64_t nq, int64_t q_per_group, const BatchScratch sc) {
  extern __shared__ __align__(16) double pl_smem[];            // [n_planes][256]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int P = ix.n_planes;
  const int64_t q0 = blockIdx.y * q_per_group;
  const int64_t q1 = min(nq, q0 + q_per_group);
  const uint32_t lt = (1u << lane) - 1u;
  const int qw = sc.qd_stride >> 3;                            // 8-byte words of packed plane slots per query
  const unsigned long long* qd64 = reinterpret_cast<const unsigned long long*>(sc.qd);
  const uint8_t* bm_bytes = reinterpret_cast<const uint8_t*>(sc.bitmap);
  double* tmax_s = pl_smem + static_cast<size_t>(P) * PLANE_TILE;                                    // [P] this tile's plane maxima
  ulonglong2* qrec = reinterpret_cast<ulonglong2*>(tmax_s + ((P + 1) & ~1));                         // [q_per_group]
  // Upper bounds: fp64 addition is monotone, so adding the plane maxima in the query's token order bounds the
  // in-order score of every doc that matches plane tokens only.  A query whose bound over the whole corpus is
  // below its threshold needs no plane pass at all; the per-tile bound prunes most (query, tile) pairs of the rest.
  for (int64_t q = q0 + threadIdx.x; q < q1; q += PL_WARPS * 32) {
    const unsigned long long t = __ldg(sc.thr + q);
    ulonglong2 rec = make_ulonglong2(0ull, ~0ull);
    if (sc.state[q] == 3) {                                    // no bound yet: this sweep fills its score histogram
      int slot = 0;
      const int n3 = min(sc.counters[2], H3_MAX);
      while (slot < n3 && sc.h3_list[slot] != static_cast<int>(q)) ++slot;
      if (slot < n3) rec = make_ulonglong2(H3_TAG | static_cast<unsigned long long>(slot), __ldg(qd64 + q * qw));
    } else if (t != THR_SKIP) {
      const double thr_d = key_f64(t);
      double ub = 0.0;
      for (int i = 0; i < sc.qd_stride; ++i) {
        const uint32_t ps = sc.qd[q * sc.qd_stride + i];
        if (ps == 0xFFu) break;
        ub = __dadd_rn(ub, __ldg(ix.plane_gmax + ps));
      }
      if (ub >= thr_d) rec = make_ulonglong2(static_cast<unsigned long long>(__double_as_longlong(thr_d)), __ldg(qd64 + q * qw));
    }
    qrec[q - q0] = rec;
  }
  // The tile of all planes is one contiguous block of P * 2 KB: one thread hands it to the bulk-copy engine in 32 KB
  // pieces and everybody waits on the mbarrier (a 512-thread LDG/STS loop needed six dependent round trips, ~9 us per
  // tile with nothing else resident on the SM to hide them).
  __shared__ __align__(8) uint64_t fill_bar;
  if (threadIdx.x == 0) { ptx::mbar_init(&fill_bar, 1); ptx::fence_barrier_init(); }
  uint32_t fill_phase = 0;
  for (int64_t tile = blockIdx.x; tile < ix.n_tiles; tile += gridDim.x) {
    __syncthreads();                                           // the previous tile is no longer read; fill_bar is initialised
    if (threadIdx.x == 0) {
      const uint32_t bytes = static_cast<uint32_t>(P) * PLANE_TILE * 8u;
      const char* src = reinterpret_cast<const char*>(ix.planes + tile * P * PLANE_TILE);
      ptx::fence_proxy_async_smem();
      ptx::mbar_expect_tx(&fill_bar, bytes);
      for (uint32_t off = 0; off < bytes; off += 32768u)
        ptx::bulk_load_1d(reinterpret_cast<char*>(pl_smem) + off, src + off, min(32768u, bytes - off), &fill_bar);
    }
    if (threadIdx.x < P) tmax_s[threadIdx.x] = __ldg(ix.plane_tmax + tile * P + threadIdx.x);
    ptx::mbar_wait(&fill_bar, fill_phase);
    fill_phase ^= 1u;
    __syncthreads();
    // Per-query inputs: {threshold, packed plane slots} come from shared memory (staged once per CTA); the member
    // byte is streamed from HBM (this tile's 32-byte records of consecutive queries).  Four queries are processed
    // per outer iteration so that the loads of the next four are in flight, in their own registers, meanwhile.
    const uint8_t* bm_tile = bm_bytes + tile * sc.nq * (PLANE_TILE / 8) + lane;   // + q * 32: this lane's 8 member bits
    auto process = [&](const int64_t q, const uint32_t mb) {
      const ulonglong2 rec = qrec[q - q0];                   // {k-th score of A as fp64 bits, packed plane slots}
      unsigned long long pk = rec.y;
        uint32_t ps = static_cast<uint32_t>(pk) & 0xFFu;
        if (ps == 0xFFu) return;                             // skipped query (no plane tokens / fallback / bounded out)
        const bool hist = (rec.x & H3_TAG) == H3_TAG;
        const double thr_d = hist ? 0.0 : __longlong_as_double(static_cast<long long>(rec.x));
        if (!hist) {
          // this tile's bound for the query (same token order); all lanes read the same shared-memory words
          double ub = tmax_s[ps];
          unsigned long long pb = pk;
          for (int i = 1; i < sc.qd_stride; ++i) {
            if ((i & 7) == 0) pb = __ldg(qd64 + q * qw + (i >> 3)); else pb >>= 8;
            const uint32_t s2 = static_cast<uint32_t>(pb) & 0xFFu;
            if (s2 == 0xFFu) break;
            ub = __dadd_rn(ub, tmax_s[s2]);
          }
          if (ub < thr_d) return;
        }
        // first plane token: 0.0 + x == x (plane values are never -0.0), so the accumulators start as its row
        double acc[8];
        {
          const double2* row = reinterpret_cast<const double2*>(pl_smem + ps * PLANE_TILE) + lane;
  #pragma unroll
          for (int j = 0; j < 4; ++j) { const double2 v = row[j * 32]; acc[2 * j] = v.x; acc[2 * j + 1] = v.y; }
        }
        for (int i = 1; i < sc.qd_stride; ++i) {
          if ((i & 7) == 0) pk = __ldg(qd64 + q * qw + (i >> 3)); else pk >>= 8;
          ps = static_cast<uint32_t>(pk) & 0xFFu;
          if (ps == 0xFFu) break;
          const double2* row = reinterpret_cast<const double2*>(pl_smem + ps * PLANE_TILE) + lane;
  #pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double2 v = row[j * 32];
            acc[2 * j] = __dadd_rn(acc[2 * j], v.x);
            acc[2 * j + 1] = __dadd_rn(acc[2 * j + 1], v.y);
          }
        }
        // docs that reach A's k-th score (a positive fp64, so comparing doubles is comparing keys) and are not
        // members of the sparse pass
        if (hist) {
          uint32_t* H = sc.hist + static_cast<size_t>(rec.x & 0xFFFFu) * H3_BINS;
  #pragma unroll
          for (int r = 0; r < 8; ++r)
            if (acc[r] > 0.0 && !((mb >> r) & 1u)) atomicAdd(&H[h3_bin(acc[r])], 1u);
          return;
        }
        uint32_t pass = 0u;
  #pragma unroll
        for (int r = 0; r < 8; ++r) if (acc[r] >= thr_d) pass |= 1u << r;
        pass &= ~mb;
        if (__ballot_sync(0xffffffffu, pass != 0u) == 0u) return;
  #pragma unroll
        for (int r = 0; r < 8; ++r) {
          const bool p = (pass >> r) & 1u;
          const uint32_t m = __ballot_sync(0xffffffffu, p);
          if (m == 0u) continue;
          int base = 0;
          if (lane == 0) base = atomicAdd(&sc.e_cnt[q], __popc(m));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (p) {
            const int pos = base + __popc(m & lt);
            if (pos < E_CAP) {
              sc.e_key[q * E_CAP + pos] = f64_key(acc[r]);
              sc.e_doc[q * E_CAP + pos] = static_cast<uint32_t>(tile * PLANE_TILE + 8 * lane + r);
            }
          }
        }
    };
    constexpr int PF = 4;
    uint32_t cur[PF], nxt[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int64_t qq = q0 + warp + static_cast<int64_t>(u) * PL_WARPS;
      cur[u] = qq < q1 ? __ldg(bm_tile + qq * 32) : 0u;
    }
    for (int64_t q = q0 + warp; q < q1; q += PF * PL_WARPS) {
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int64_t qq = q + static_cast<int64_t>(PF + u) * PL_WARPS;
        nxt[u] = qq < q1 ? __ldg(bm_tile + qq * 32) : 0u;
      }
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int64_t qq = q + static_cast<int64_t>(u) * PL_WARPS;
        if (qq < q1) process(qq, cur[u]);
      }
#pragma unroll
      for (int u = 0; u < PF; ++u) cur[u] = nxt[u];
    }
  }
}

// The histogram path costs one extra sweep over all plane tiles; the per-query fallback costs ~13 us per query and
// 16K-doc block.  With fewer than H3_MIN listed queries the fallback is cheaper: hand them over.  One block.
constexpr int H3_MIN = 64;
__global__ void __launch_bounds__(256)
bm25_h3_decide_kernel(const BatchScratch sc) {
  const int n3 = min(sc.counters[2], H3_MAX);
  if (n3 == 0 || n3 >= H3_MIN) return;
  for (int i = threadIdx.x; i < n3; i += blockDim.x) {
    const int q = sc.h3_list[i];
    sc.state[q] = 2;
    sc.flag_list[atomicAdd(&sc.counters[1], 1)] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) sc.counters[2] = 0;
}

// ----------------------------------------------------------------------------
// 2b. plane pass over the (few) queries of h3_list: emission against the threshold that bm25_h3_threshold_kernel
// derived from the histogram the main sweep filled.  Returns at once when the list is empty.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(PL_WARPS * 32, 1)
bm25_plane_list_kernel(const Bm25Dev ix, const BatchScratch sc) {
  extern __shared__ __align__(16) double pll_smem[];           // [n_planes][256]
  const int n3 = min(sc.counters[2], H3_MAX);
  if (n3 == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int P = ix.n_planes;
  const uint32_t lt = (1u << lane) - 1u;
  const uint8_t* bm_bytes = reinterpret_cast<const uint8_t*>(sc.bitmap);
  for (int64_t tile = blockIdx.x; tile < ix.n_tiles; tile += gridDim.x) {
    __syncthreads();
    {
      const double2* src = reinterpret_cast<const double2*>(ix.planes + tile * P * PLANE_TILE);
      double2* dst = reinterpret_cast<double2*>(pll_smem);
      for (int i = threadIdx.x; i < P * (PLANE_TILE / 2); i += PL_WARPS * 32) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    for (int idx = warp; idx < n3; idx += PL_WARPS) {
      const int64_t q = sc.h3_list[idx];
      const unsigned long long t = sc.thr[q];
      if (t == THR_SKIP) continue;
      const double thr_d = key_f64(t);
      double acc[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] = 0.0;
      for (int i = 0; i < sc.qd_stride; ++i) {
        const uint32_t ps = sc.qd[q * sc.qd_stride + i];
        if (ps == 0xFFu) break;
        const double2* row = reinterpret_cast<const double2*>(pll_smem + ps * PLANE_TILE) + lane;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double2 v = row[j * 32];
          acc[2 * j] = __dadd_rn(acc[2 * j], v.x);
          acc[2 * j + 1] = __dadd_rn(acc[2 * j + 1], v.y);
        }
      }
      const uint32_t mb = bm_bytes[(tile * sc.nq + q) * (PLANE_TILE / 8) + lane];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const bool p = acc[r] > 0.0 && !((mb >> r) & 1u) && acc[r] >= thr_d;
        const uint32_t m = __ballot_sync(0xffffffffu, p);
        if (m == 0u) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(&sc.e_cnt[q], __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (p) {
          const int pos = base + __popc(m & lt);
          if (pos < E_CAP) {
            sc.e_key[q * E_CAP + pos] = f64_key(acc[r]);
            sc.e_doc[q * E_CAP + pos] = static_cast<uint32_t>(tile * PLANE_TILE + 8 * lane + r);
          }
        }
      }
    }
  }
}

// threshold = lower edge of the highest bin at which the count of scores at or above it reaches k
// (every positive score when there are fewer); one warp per listed query
__global__ void __launch_bounds__(128)
bm25_h3_threshold_kernel(int k, const BatchScratch sc) {
  const int n3 = min(sc.counters[2], H3_MAX);
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (idx >= n3) return;
  const int64_t q = sc.h3_list[idx];
  const int need = k;        // the members of A may all rank below the k best non-members
  const uint32_t* H = sc.hist + static_cast<size_t>(idx) * H3_BINS;
  int running = 0, bin = 0;
  for (int c = H3_BINS / 32 - 1; c >= 0; --c) {
    const int h = static_cast<int>(H[c * 32 + lane]);
    int suf = h;                                               // suffix sum over lanes >= lane
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += v; }
    const uint32_t reach = __ballot_sync(0xffffffffu, running + suf >= need);
    if (reach) { bin = c * 32 + (31 - __clz(reach)); break; }
    running += __shfl_sync(0xffffffffu, suf, 0);
  }
  if (lane == 0) { sc.thr[q] = h3_edge(bin); sc.state[q] = 0; }
}

// ----------------------------------------------------------------------------
// 3. final merge: one CTA per query.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
bm25_final_merge_kernel(int64_t nq, int k, int64_t id_base, const BatchScratch sc, double* __restrict__ out_scores,
                        int64_t* __restrict__ out_ids) {
  extern __shared__ __align__(16) uint8_t fm_smem[];
  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x;
  const int st = sc.state[q];
  if (st == 2) return;                                         // the fallback writes this row
  const int ne = st == 0 ? sc.e_cnt[q] : 0;
  if (ne > E_CAP) {                                            // emissions overflowed (mass ties): fallback
    if (tid == 0) { sc.state[q] = 2; sc.flag_list[atomicAdd(&sc.counters[1], 1)] = static_cast<int>(q); }
    return;
  }
  const int na = sc.a_cnt[q];
  if (ne == 0) {
    for (int i = tid; i < k; i += blockDim.x) {
      const bool ok = i < na;
      out_scores[q * k + i] = ok ? key_f64(sc.a_key[q * k + i]) : 0.0;
      out_ids[q * k + i] = ok ? id_base + static_cast<int64_t>(sc.a_doc[q * k + i]) : -1;
    }
    return;
  }
  const int n = na + ne;
  int n2 = 2;
  while (n2 < n) n2 <<= 1;
  unsigned long long* key = reinterpret_cast<unsigned long long*>(fm_smem);
  uint32_t* doc = reinterpret_cast<uint32_t*>(fm_smem + static_cast<size_t>(ARB_MAX_TOPK + E_CAP) * 2 * 8);
  for (int i = tid; i < n2; i += blockDim.x) {
    unsigned long long kk = 0ull; uint32_t dd = 0xFFFFFFFFu;
    if (i < na) { kk = sc.a_key[q * k + i]; dd = sc.a_doc[q * k + i]; }
    else if (i < n) { kk = sc.e_key[q * E_CAP + (i - na)]; dd = sc.e_doc[q * E_CAP + (i - na)]; }
    key[i] = kk; doc[i] = dd;
  }
  cta_sort_desc(key, doc, n2);
  for (int i = tid; i < k; i += blockDim.x) {
    const bool ok = i < n;
    out_scores[q * k + i] = ok ? key_f64(key[i]) : 0.0;
    out_ids[q * k + i] = ok ? id_base + static_cast<int64_t>(doc[i]) : -1;
  }
}
constexpr int FM_SMEM = (ARB_MAX_TOPK + E_CAP) * 2 * 12;

// merge the per-split part lists of the fallback for the flagged queries
__global__ void __launch_bounds__(256)
bm25_flag_merge_kernel(const double* __restrict__ ps, const int64_t* __restrict__ pi, int n_split, int64_t nq, int k,
                       const int* __restrict__ flag_list, const int* __restrict__ flag_count, int n_pow2,
                       double* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  extern __shared__ __align__(16) uint8_t gm_smem[];
  unsigned long long* key = reinterpret_cast<unsigned long long*>(gm_smem);
  int64_t* id = reinterpret_cast<int64_t*>(gm_smem + static_cast<size_t>(n_pow2) * 8);
  const int n_listed = *flag_count;
  const int m = n_split * k;
  for (int slot = blockIdx.x; slot < n_listed; slot += gridDim.x) {
    const int64_t q = flag_list[slot];
    __syncthreads();
    for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
      unsigned long long kk = 0ull; int64_t ii = 0x7fffffffffffffffll;
      if (i < m) {
        const int64_t o = (static_cast<int64_t>(i / k) * nq + q) * k + (i % k);
        const int64_t v = pi[o];
        if (v >= 0) { ii = v; kk = f64_key(ps[o]); }
      }
      key[i] = kk; id[i] = ii;
    }
    cta_sort_desc(key, id, n_pow2);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      const bool ok = i < n_pow2 && key[i] != 0ull;
      out_scores[q * k + i] = ok ? key_f64(key[i]) : 0.0;
      out_ids[q * k + i] = ok ? id[i] : -1;
    }
  }
}

static thread_local int64_t g_stats[4] = {0, 0, 0, 0};

static int batch_set_attrs() {
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    err = cudaFuncSetAttribute(bm25_sparse_pass_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               sp_warps<256>() * static_cast<int>(sizeof(SpWarp<256>)));
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(bm25_sparse_pass_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 sp_warps<512>() * static_cast<int>(sizeof(SpWarp<512>)));
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(bm25_plane_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PL_SMEM_MAX);
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(bm25_plane_list_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_PLANES * PLANE_TILE * 8);
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(bm25_flag_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16);
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(bm25_a_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12);
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(bm25_a_merge_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 AMW_WARPS * static_cast<int>(sizeof(AmWarp)));

  });
  ARB_CUDA(err);
  return bm25_set_smem_attrs();
}

static int batch_chunk(const Bm25Dev& ix, const int32_t* d_terms, int64_t nq, int mq, int k, int64_t id_base,
                       double* d_scores, int64_t* d_ids, cudaStream_t st) {
  const int sms = num_sms();
  const bool planes = ix.n_planes > 0;
  BatchScratch sc{};
  sc.bm_tiles = static_cast<int64_t>(ix.nb) * (DOC_BLOCK / PLANE_TILE);  // whole 16K-doc blocks
  sc.nq = nq;
  AsyncBuf b_akey, b_adoc, b_i32, b_thr, b_u8, b_bitmap, b_ekey, b_edoc, b_pakey, b_padoc, b_pacnt, b_gthr, b_hist;
  sc.parts = static_cast<int>(std::min<int64_t>(ix.nb, 24));
  if (const char* v = std::getenv("ARB_BM25_PARTS")) sc.parts = std::max(1, std::min<int>(std::min<int64_t>(ix.nb, 32), atoi(v)));
  ARB_CUDA(b_pakey.alloc(static_cast<size_t>(nq) * sc.parts * k * 8, st));
  ARB_CUDA(b_padoc.alloc(static_cast<size_t>(nq) * sc.parts * k * 4, st));
  ARB_CUDA(b_pacnt.alloc(static_cast<size_t>(nq) * sc.parts * 4, st));
  ARB_CUDA(b_gthr.alloc(static_cast<size_t>(nq) * 8, st));
  ARB_CUDA(cudaMemsetAsync(b_gthr.p, 0, static_cast<size_t>(nq) * 8, st));
  sc.gthr = b_gthr.as<unsigned long long>();
  sc.pa_key = b_pakey.as<unsigned long long>();
  sc.pa_doc = b_padoc.as<uint32_t>();
  sc.pa_cnt = b_pacnt.as<int>();
  ARB_CUDA(b_akey.alloc(static_cast<size_t>(nq) * k * 8, st));
  ARB_CUDA(b_adoc.alloc(static_cast<size_t>(nq) * k * 4, st));
  ARB_CUDA(b_i32.alloc((static_cast<size_t>(nq) * 3 + 4 + H3_MAX) * 4, st));   // e_cnt | counters | a_cnt | flag_list | h3_list
  ARB_CUDA(b_thr.alloc(static_cast<size_t>(nq) * 8, st));
  sc.qd_stride = (mq + 7) & ~7;
  ARB_CUDA(b_u8.alloc(static_cast<size_t>(nq) * (sc.qd_stride + 8) + 16, st));
  if (planes) {
    ARB_CUDA(b_bitmap.alloc(static_cast<size_t>(nq) * sc.bm_tiles * 32, st));
    ARB_CUDA(b_ekey.alloc(static_cast<size_t>(nq) * E_CAP * 8, st));
    ARB_CUDA(b_edoc.alloc(static_cast<size_t>(nq) * E_CAP * 4, st));
  }
  sc.a_key = b_akey.as<unsigned long long>();
  sc.a_doc = b_adoc.as<uint32_t>();
  sc.e_cnt = b_i32.as<int>();
  sc.counters = sc.e_cnt + nq;
  sc.a_cnt = sc.counters + 4;
  sc.flag_list = sc.a_cnt + nq;
  sc.h3_list = sc.flag_list + nq;
  sc.thr = b_thr.as<unsigned long long>();
  sc.qd = b_u8.as<uint8_t>();                                   // 8-byte aligned rows first
  sc.state = sc.qd + static_cast<size_t>(nq) * sc.qd_stride;
  sc.bitmap = planes ? b_bitmap.as<uint32_t>() : nullptr;
  sc.e_key = planes ? b_ekey.as<unsigned long long>() : nullptr;
  sc.e_doc = planes ? b_edoc.as<uint32_t>() : nullptr;
  ARB_CUDA(cudaMemsetAsync(sc.e_cnt, 0, (static_cast<size_t>(nq) + 4) * 4, st));
  if (planes) {
    ARB_CUDA(b_hist.alloc(static_cast<size_t>(H3_MAX) * H3_BINS * 4, st));
    sc.hist = b_hist.as<uint32_t>();
    ARB_CUDA(cudaMemsetAsync(sc.hist, 0, static_cast<size_t>(H3_MAX) * H3_BINS * 4, st));
  }

  {
    ProfileScope ps(PROF_BM25_SPARSE, st);
    if (2 * k + 32 <= 256) {
      constexpr int W = sp_warps<256>();
      const unsigned grid = static_cast<unsigned>(std::min<int64_t>(sms, (nq * sc.parts + W - 1) / W));
      bm25_sparse_pass_kernel<256><<<grid, W * 32, W * sizeof(SpWarp<256>), st>>>(ix, d_terms, mq, nq, k, sc);
    } else {
      constexpr int W = sp_warps<512>();
      const unsigned grid = static_cast<unsigned>(std::min<int64_t>(sms, (nq * sc.parts + W - 1) / W));
      bm25_sparse_pass_kernel<512><<<grid, W * 32, W * sizeof(SpWarp<512>), st>>>(ix, d_terms, mq, nq, k, sc);
    }
    ARB_LAUNCH_OK();
    if (sc.parts * k <= AMW_KEYS) {
      bm25_a_merge_warp_kernel<<<static_cast<unsigned>((nq + AMW_WARPS - 1) / AMW_WARPS), AMW_WARPS * 32,
                                 AMW_WARPS * sizeof(AmWarp), st>>>(nq, k, sc);
    } else {
      int n2 = 2;
      while (n2 < sc.parts * k) n2 <<= 1;
      bm25_a_merge_kernel<<<static_cast<unsigned>(nq), 128, static_cast<size_t>(n2) * 12, st>>>(nq, k, n2, sc);
    }
    ARB_LAUNCH_OK();
  }
  if (planes) {
    ProfileScope ps(PROF_BM25_PLANES, st);
    bm25_h3_decide_kernel<<<1, 256, 0, st>>>(sc);
    ARB_LAUNCH_OK();
    // enough CTAs to fill the GPU: tile sets first, query groups when the corpus has few tiles
    const int64_t sets = std::min<int64_t>(ix.n_tiles, sms);
    const size_t plane_smem = static_cast<size_t>(ix.n_planes) * PLANE_TILE * 8;
    const size_t tmax_smem = static_cast<size_t>((ix.n_planes + 1) & ~1) * 8;
    const int64_t max_qpg = static_cast<int64_t>((PL_SMEM_MAX - plane_smem - tmax_smem) / 16);   // {thr, slots} records per CTA
    int64_t groups = std::max<int64_t>(1, std::min<int64_t>(sms / sets, (nq + 4 * PL_WARPS - 1) / (4 * PL_WARPS)));
    groups = std::max<int64_t>(groups, (nq + max_qpg - 1) / max_qpg);
    const int64_t qpg = (nq + groups - 1) / groups;
    groups = (nq + qpg - 1) / qpg;
    dim3 grid(static_cast<unsigned>(sets), static_cast<unsigned>(groups));
    bm25_plane_pass_kernel<<<grid, PL_WARPS * 32, plane_smem + tmax_smem + static_cast<size_t>(qpg) * 16, st>>>(ix, nq, qpg, sc);
    ARB_LAUNCH_OK();
    // queries without a bound from list A: their score histogram was filled by the sweep above -> threshold -> emission
    bm25_h3_threshold_kernel<<<H3_MAX / 4, 128, 0, st>>>(k, sc);
    ARB_LAUNCH_OK();
    bm25_plane_list_kernel<<<static_cast<unsigned>(sets), PL_WARPS * 32, plane_smem, st>>>(ix, sc);
    ARB_LAUNCH_OK();
  }
  {
    ProfileScope ps(PROF_BM25_FINAL, st);
    bm25_final_merge_kernel<<<static_cast<unsigned>(nq), 128, FM_SMEM, st>>>(nq, k, id_base, sc, d_scores, d_ids);
    ARB_LAUNCH_OK();
    {
      // fallback for the flagged queries (the count lives on the device; CTAs with nothing to do exit)
      // (8 CTAs per query: 32 measured slower at 4096 queries -- 0.49 vs 0.29 ms -- the partial lists cost more than the latency saved)
      int n_split = static_cast<int>(std::min<int64_t>(ix.nb, 8));
      n_split = std::max(1, std::min(n_split, 4096 / k));
      const int bps = (ix.nb + n_split - 1) / n_split;
      n_split = (ix.nb + bps - 1) / bps;
      const int grid_x = static_cast<int>(std::min<int64_t>(nq, std::max(1, 2 * sms / n_split)));
      AsyncBuf pscores, pids;
      double* ps_ = d_scores;
      int64_t* pi_ = d_ids;
      if (n_split > 1) {
        ARB_CUDA(pscores.alloc(static_cast<size_t>(n_split) * nq * k * 8, st));
        ARB_CUDA(pids.alloc(static_cast<size_t>(n_split) * nq * k * 8, st));
        ps_ = pscores.as<double>();
        pi_ = pids.as<int64_t>();
      }
      int rc = launch_bm25_topk(ix, d_terms, mq, k, id_base, nq, n_split, bps, sc.flag_list, sc.counters + 1, grid_x,
                                ps_, pi_, st);
      if (rc) return rc;
      if (n_split > 1) {
        int n_pow2 = 2;
        while (n_pow2 < n_split * k) n_pow2 <<= 1;
        bm25_flag_merge_kernel<<<static_cast<unsigned>(grid_x), 256, static_cast<size_t>(n_pow2) * 16, st>>>(
            ps_, pi_, n_split, nq, k, sc.flag_list, sc.counters + 1, n_pow2, d_scores, d_ids);
        ARB_LAUNCH_OK();
      }
    }
  }
  if (std::getenv("ARB_BM25_STATS")) {                           // diagnostics only: synchronises
    std::vector<int> h(static_cast<size_t>(nq) + 4);
    ARB_CUDA(cudaMemcpyAsync(h.data(), sc.e_cnt, h.size() * 4, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaStreamSynchronize(st));
    g_stats[0] += nq;
    g_stats[1] += h[nq + 1];
    for (int64_t i = 0; i < nq; ++i) { g_stats[2] += h[i]; g_stats[3] = std::max<int64_t>(g_stats[3], h[i]); }
  }
  return ARB_OK;
}

int bm25_batch_search(const Bm25Dev& ix, const int32_t* d_terms, int64_t nq, int max_qt, int k, int64_t id_base,
                      double* d_scores, int64_t* d_ids, cudaStream_t st) {
  int rc = batch_set_attrs();
  if (rc) return rc;
  // the member bitmap is nq x n_docs bits: cut the batch so that it stays under ~2 GB
  const int64_t row_bytes = std::max<int64_t>(1, static_cast<int64_t>(ix.nb) * (DOC_BLOCK / 8));
  int64_t chunk = std::max<int64_t>(64, (int64_t(2) << 30) / row_bytes);
  if (const char* v = std::getenv("ARB_BM25_CHUNK")) chunk = std::max<int64_t>(1, atoll(v));
  for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
    const int64_t n = std::min(chunk, nq - q0);
    rc = batch_chunk(ix, d_terms + q0 * max_qt, n, max_qt, k, id_base, d_scores + q0 * k, d_ids + q0 * k, st);
    if (rc) return rc;
  }
  return ARB_OK;
}

}  // namespace arb

extern "C" int arb_bm25_batch_stats(int64_t out[4], int reset) {
  for (int i = 0; i < 4; ++i) { if (out) out[i] = arb::g_stats[i]; if (reset) arb::g_stats[i] = 0; }
  return ARB_OK;
}
