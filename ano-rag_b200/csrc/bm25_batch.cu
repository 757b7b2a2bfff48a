// Batched BM25 top-k (sm_100a): the tile-major join pipeline behind arb_bm25_search.
//
// Replaces corpus.get_scores() + top-k for a whole query batch (reference
// utils/bm25_search.py:309 -> rank_bm25 BM25Okapi.get_scores / SimpleBM25.get_scores :43-63).
//
// In a Zipf corpus a few dozen terms own almost all postings (cfg3: the 96 densest terms carry 98 % of the
// postings a query touches).  Every such term is stored as a doc-major fp64 "plane" of its finished
// contributions idf*g (+0.0 where the term is absent); all other ("sparse") postings are stored twice: term-major
// for the per-query kernel and doc-major ("doc rows": per doc its sparse terms, ascending, with the finished
// contribution).  A batch is scored in doc order, 256-doc tiles at a time, where the tile of all planes sits in
// shared memory:
//
//   0. query-term index   the batch's sparse query tokens grouped by term: bt_off[term] .. -> (query, position).
//   1. member pass        every doc row of a tile is JOINED with that index: each (doc, sparse term) posting
//      whose term occurs in the batch yields the (query, position) records that match it, one pair per thread
//      (counts -> CTA-wide scan -> 512 independent pairs in flight).  The thread computes the COMPLETE score of
//      its (query, doc) in the reference's order -- query tokens first to last, plane tokens from the staged
//      tile, the query's other sparse tokens looked up in the doc's own row (a 384-bit two-hash Bloom filter of
//      the row answers almost all of those look-ups; positives are deferred and scored together) -- and, when
//      another sparse token of the query matches the doc at an earlier position, leaves the pair to that thread.  Scores
//      above the query's running bound are appended to its candidate list.  The corpus is swept in a few rounds
//      of growing size; between rounds a compaction kernel cuts every list back to its k best and publishes the
//      k-th score as the bound of the next round (rounds go in ascending doc order, so a later doc that only
//      ties the bound loses the tie).  The last compaction leaves list A = top-k of the docs that match at
//      least one sparse token ("members").
//   2. screening pass     every other doc has plane tokens only.  fp32 copies of the planes, rounded up, are
//      summed with upward rounding: a rigorous upper bound of the in-order fp64 sum at half the shared-memory
//      traffic.  Only (query, tile) pairs in which some bound reaches A's k-th score are recomputed in fp64
//      (same token order); docs that reach it and are not members of the query go to list E.
//   3. final merge        A + E -> top-k by (score desc, doc asc).
//   4. fallback           queries whose A holds fewer than k docs while they have plane tokens take a histogram
//      pass (threshold from the distribution of plane-only scores) or, like queries whose lists overflowed,
//      the per-query kernel (bm25.cu).
//
// Adding +0.0 for absent terms is what the reference itself does (score += idf * 0), and an accumulator that
// starts at +0.0 can never become -0.0, so the result is bit-identical.
#include "bm25.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <vector>

namespace arb {

constexpr int SC_GRAB = 8;                           // active queries a warp takes at a time: small grabs keep the 32 warps of a CTA
                                                     // busy until the tile's barrier (32 per grab left a third of the last wave idle)
constexpr int SC_WARPS = 32;                         // screening pass: warps per CTA (two fp32 plane tiles in flight)
constexpr int PL_WARPS = 16;                         // fp64 plane kernels of the histogram path
constexpr int SMEM_MAX = 227 * 1024 - 256;           // dynamic shared memory per CTA (256 B left for static barriers / scan scratch)
constexpr int C_CAP = 4096;                          // candidate entries a query may collect between two compactions
constexpr int E_CAP = 1024;                          // screening-pass emissions kept per query
constexpr int RQ_CAP = 1 << 20;                      // queued fp64 rechecks per batch chunk (more: done in place)
constexpr unsigned long long THR_SKIP = ~0ull;
constexpr int H3_MAX = 1024;                         // queries per batch chunk that may take the histogram path
constexpr int H3_DIRECT = 64;                        // up to this many listed queries read their plane rows straight from global memory
constexpr int H3_BINS = 4096;                        // 128 bins per octave from 2^-12 up
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

struct MjDefer;
struct BatchScratch {
  int64_t nq;
  // ---- query-term index of the batch
  int32_t* qinfo;              // [nq, qs] per query token: >= 0 sparse term id, -1 contributes nothing, <= -2 plane slot (-2 - slot)
  int qs;                      // tokens per query rounded up to a multiple of 4 (16-byte rows)
  uint32_t* qdup;              // [nq] bit p: token p repeats a sparse term of an earlier position (it gets no record of its own)
  uint2* bt_ent;               // [n_terms] {first record, number of records} of a term in bt_rec (count 0: not in the batch)
  uint32_t* bt_cur;            // [n_terms] 0 = unclaimed, then 1 + records placed so far
  int4* bt_rec;                // [sparse tokens of the batch] records of 1 + qs/4 int4, grouped by term (bm25_bt_fill_kernel)
  uint32_t* rec_of;            // [nq * qs] record of a sparse query token, else 0xFFFFFFFF
  MjDefer* defer;
  int* def_cnt;                // [member-pass CTAs] entries queued by the launch       // [member-pass CTAs, MJ_DEF] set-aside pairs of the unit being joined
  // ---- member pass
  unsigned long long* c_key;   // [nq, C_CAP] candidates; [0, k) = running top-k after a compaction
  uint32_t* c_doc;             // [nq, C_CAP]
  int* c_cnt;                  // [nq] entries appended (may exceed C_CAP: overflow)
  unsigned long long* gthr;    // [nq] bound of the current round: only scores strictly above it are candidates
  int* n_nonpos;               // [nq] members whose complete score is not positive (never candidates)
  uint8_t* ovf;                // [nq] a list overflowed: the per-query kernel takes the query
  // ---- list A and the per-query state the later passes read
  unsigned long long* a_key;   // [nq, k] top-k of the members, sorted
  uint32_t* a_doc;             // [nq, k]
  int* a_cnt;                  // [nq]
  unsigned long long* thr;     // [nq]  key of A's k-th score, or THR_SKIP
  uint8_t* state;              // [nq]  0 = screening pass runs, 1 = no plane tokens (A is final), 2 = fallback, 3 = histogram path
  uint8_t* qd;                 // [nq, qd_stride] the query's plane slots in token order, 0xFF terminated/padded
  int qd_stride;               // mq rounded up to a multiple of 8 (<= 32)
  unsigned long long* e_key;   // [nq, E_CAP]
  uint32_t* e_doc;             // [nq, E_CAP]
  int* e_cnt;                  // [nq]
  int* flag_list;              // [nq]
  int* counters;               // [0] set-aside pairs, [1] flagged queries, [2] histogram queries, [3] queued rechecks, [4] set-aside pairs
                               // that did match another sparse token (diagnostics), [5] records handed out
  int stats_on;
  uint2* rq;                   // [RQ_CAP] (query, tile) pairs the screening pass wants recomputed in fp64
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


constexpr uint32_t SLOT_NONE = 127u;
constexpr int MJ_ROW_BYTES = (PLANE_TILE + 4) * 4;     // padded fp32 plane rows of the member pass: 1040 B keeps bulk copies 16-byte aligned
__host__ __device__ __forceinline__ int rec_int4s(int qs) { return 2 + (qs >> 1); }   // int4 words of a query-term record
constexpr int MJ_BLOOM_W = BLOOM_W;                  // bm25.cuh: the per-doc filters are part of the index

// ----------------------------------------------------------------------------
// 0. query-term index of the batch
// ----------------------------------------------------------------------------
// one thread per query: classify its tokens, count the sparse ones per term, pack its plane slots
__global__ void __launch_bounds__(128)
bm25_qclassify_kernel(const Bm25Dev ix, const int32_t* __restrict__ q_terms, int mq, int64_t nq, const BatchScratch sc) {
  const int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (q >= nq) return;
  int np = 0;
  uint32_t dup = 0u;
  for (int p = 0; p < sc.qs; ++p) {
    int code = -1;
    const int t = p < mq ? q_terms[q * mq + p] : -1;
    if (t >= 0 && t < ix.n_terms && __ldg(ix.idf + t) != 0.0) {        // (idf.get(q) or 0) == 0 adds +0.0: a no-op
      const int ps = ix.n_planes > 0 ? __ldg(ix.plane_slot + t) : -1;
      if (ps >= 0) { code = -2 - ps; sc.qd[q * sc.qd_stride + np++] = static_cast<uint8_t>(ps); }
      else {
        // a term that occurs twice in a query adds its contribution twice (the reference sums per token); the first
        // occurrence's record carries the repeats, so that a doc holding the term is joined once, not once per repeat
        code = t;
        bool rep = false;
        for (int e = 0; e < p; ++e) rep |= sc.qinfo[q * sc.qs + e] == t;
        if (rep) dup |= 1u << p; else atomicAdd(&sc.bt_ent[t].y, 1u);
      }
    }
    sc.qinfo[q * sc.qs + p] = code;
  }
  sc.qdup[q] = dup;
  for (int i = np; i < sc.qd_stride; ++i) sc.qd[q * sc.qd_stride + i] = 0xFF;
  sc.gthr[q] = f64_key(0.0);                                            // only positive scores are candidates
}

// The records of a term are contiguous; their place is handed out by whichever token of the term gets there first
// (no pass over the vocabulary: only its per-batch zeroing is O(n_terms)).
__global__ void __launch_bounds__(256)
bm25_bt_alloc_kernel(int64_t n, const BatchScratch sc) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int code = sc.qinfo[i];
  if (code < 0 || ((sc.qdup[i / sc.qs] >> (i % sc.qs)) & 1u)) return;
  if (atomicExch(sc.bt_cur + code, 1u) == 0u)
    sc.bt_ent[code].x = static_cast<uint32_t>(atomicAdd(&sc.counters[5], static_cast<int>(sc.bt_ent[code].y)));
}

// One record per sparse query token, grouped by term: everything the member pass needs to score the token's query
// against a doc arrives DECODED -- int4 #0 {query << 5 | position, bitmap of the plane tokens, bound of the current round
// (8 B)}, int4 #1 {bitmap of the query's OTHER sparse tokens}, then 8 bytes per query token: {byte offset of its plane
// row in the staged tile (bits 0-16; the zero row when it is no plane term) | byte offset of its filter word (17-22) |
// plane slot (23-29, 127 = none), filter mask (one to three bits; bit 31, never set in a filter, for tokens that are not
// another sparse term)}.  The member pass starts a new record every few pairs, with a third of its lanes: decoding there
// cost as many issue slots as scoring.  rec_of remembers where a query's records went (the compaction updates their bound).
__global__ void __launch_bounds__(256)
bm25_bt_fill_kernel(int64_t n, const BatchScratch sc, int n_planes) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int code = sc.qinfo[i];
  const int64_t q = i / sc.qs;
  const int own = static_cast<int>(i - q * sc.qs);
  if (code < 0 || ((sc.qdup[q] >> own) & 1u)) { sc.rec_of[i] = 0xFFFFFFFFu; return; }
  const uint32_t at = sc.bt_ent[code].x + atomicAdd(sc.bt_cur + code, 1u) - 1u;
  sc.rec_of[i] = at;
  int4* rec = sc.bt_rec + static_cast<size_t>(at) * rec_int4s(sc.qs);
  const int* src = sc.qinfo + q * sc.qs;
  uint32_t plane_bits = 0u, sparse_bits = 0u, own_bits = 1u << own;
  for (int j = 0; j < (sc.qs >> 1); ++j) {
    uint32_t f[4];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int p = 2 * j + e;
      const int c = src[p];
      uint32_t slot = SLOT_NONE, po = static_cast<uint32_t>(n_planes) * MJ_ROW_BYTES, bw = 0u, bm = 1u << 31;
      if (c <= -2) { slot = static_cast<uint32_t>(-2 - c); po = slot * MJ_ROW_BYTES; plane_bits |= 1u << p; }
      else if (c == code) own_bits |= 1u << p;                                   // the own token and its repeats
      else if (c >= 0) {
        bm = bloom_mask(static_cast<uint32_t>(c));                             // one to three distinct bits
        bw = bloom_word(static_cast<uint32_t>(c)) * 4u;
        sparse_bits |= 1u << p;
      }
      f[2 * e] = po | (bw << 17) | (slot << 23);
      f[2 * e + 1] = bm;
    }
    rec[2 + j] = make_int4(static_cast<int>(f[0]), static_cast<int>(f[1]), static_cast<int>(f[2]), static_cast<int>(f[3]));
  }
  const unsigned long long b0 = f64_key(0.0);
  rec[0] = make_int4(static_cast<int>(static_cast<uint32_t>(q << 5) | static_cast<uint32_t>(own)), static_cast<int>(plane_bits),
                     static_cast<int>(static_cast<uint32_t>(b0)), static_cast<int>(static_cast<uint32_t>(b0 >> 32)));
  rec[1] = make_int4(static_cast<int>(sparse_bits), static_cast<int>(own_bits), 0, 0);
}

// ----------------------------------------------------------------------------
// 1. member pass
// ----------------------------------------------------------------------------
constexpr int MJ_THREADS = 768;
constexpr int MJ_PPT = 4;                            // postings a thread owns in the join's preparation
constexpr int MJ_CH_MAX = MJ_THREADS * MJ_PPT;       // postings joined per chunk
constexpr int MJ_LAUNCH_TILES = 2048;                 // tiles per member-pass launch
constexpr int MJ_DEF = 65536;                        // set-aside pairs per CTA and launch, in global scratch (more: finished in place)
constexpr int MJ_PSTRIDE = MJ_ROW_BYTES / 4;
// dynamic shared memory of the member pass: [P + 1][MJ_PSTRIDE] fp32 plane rows, rounded up (row P = zeros) |
// bloom [256][16] | pscan [ch + 4] | cbase [ch] | pinfo [ch]
__host__ __device__ inline size_t mj_fixed_bytes(int P) { return static_cast<size_t>(P + 1) * MJ_PSTRIDE * 4 + PLANE_TILE * MJ_BLOOM_W * 4 + 16; }
// index of term c in the ascending list dt[0, n), or -1
__device__ __forceinline__ int row_find(const uint32_t* __restrict__ dt, int n, uint32_t c) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(dt + mid) < c) lo = mid + 1; else hi = mid; }
  return (lo < n && __ldg(dt + lo) == c) ? lo : -1;
}
// does doc d match one of query q's sparse tokens?  (rare path: screening emissions only)
__device__ __forceinline__ bool is_member(const uint32_t* __restrict__ drow_off, const uint32_t* __restrict__ drow_term,
                                          const int32_t* __restrict__ qinfo, int qs, int64_t q, int64_t d) {
  const uint32_t o0 = __ldg(drow_off + d);
  const int n = static_cast<int>(__ldg(drow_off + d + 1) - o0);
  if (n == 0) return false;
  for (int p = 0; p < qs; ++p) {
    const int c = __ldg(qinfo + q * qs + p);
    if (c >= 0 && row_find(drow_term + o0, n, static_cast<uint32_t>(c)) >= 0) return true;
  }
  return false;
}
// largest i in [0, n) with off[i] <= x (off ascending, off[0] <= x < off[n])
__device__ __forceinline__ int seg_of(const uint32_t* off, int n, uint32_t x) {
  int lo = 0, hi = n;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= x) lo = mid; else hi = mid; }
  return lo;
}

// append (key, doc) to query q's candidates if it beats the round's bound; count members that can never be candidates
template <class Scratch>
__device__ __forceinline__ void member_emit(const Scratch& sc, int64_t q, double acc, unsigned long long bound, int64_t d) {
  const unsigned long long key = f64_key(acc);
  if (key > bound) {
    const int at = atomicAdd(sc.c_cnt + q, 1);
    if (at < C_CAP) { sc.c_key[q * C_CAP + at] = key; sc.c_doc[q * C_CAP + at] = static_cast<uint32_t>(d); }
  } else if (!(acc > 0.0)) {
    atomicAdd(sc.n_nonpos + q, 1);
  }
}

// One set-aside (record, doc) pair: the same in-order sum as the fast path, but the positions whose filter look-up was
// positive are looked up in the doc's row (binary search).  A pair whose query matches the doc at an EARLIER position
// belongs to that position's pair and is dropped.  `plane(slot)` returns the doc's value in plane `slot`.
struct ExactArgs {
  const uint32_t* drow_off; const uint32_t* drow_term; const double* drow_val; const int32_t* qinfo;
  unsigned long long* c_key; uint32_t* c_doc; int* c_cnt; int* n_nonpos; int qs, n_planes; int* tp_count;
};
template <int QS, class PlaneFn>
__device__ __forceinline__ void exact_pair(const ExactArgs& ea, PlaneFn plane, int64_t d, double myval, const int4* __restrict__ rec,
                                           uint32_t hits) {
  const int4 hd = __ldg(rec);
  const int64_t q = static_cast<uint32_t>(hd.x) >> 5;
  const int pos = hd.x & 31;
  const uint32_t own_bits = static_cast<uint32_t>(__ldg(reinterpret_cast<const int*>(rec + 1) + 1));
  const uint32_t r0 = __ldg(ea.drow_off + d);
  const int row_n = static_cast<int>(__ldg(ea.drow_off + d + 1) - r0);
  // a match at an earlier position: the pair belongs to that position's record -- settled before any plane is read
  for (uint32_t early = hits & ((1u << pos) - 1u); early != 0u; early &= early - 1u) {
    const int pp = __ffs(early) - 1;
    if (row_find(ea.drow_term + r0, row_n, static_cast<uint32_t>(__ldg(ea.qinfo + q * QS + pp))) >= 0) {
      if (ea.tp_count) atomicAdd(ea.tp_count, 1);              // diagnostics
      return;
    }
  }
  hits &= ~((2u << pos) - 1u);                                 // earlier positions did not match: they add +0.0
  // every plane value of the pair is requested before the first one is used (independent addresses: the loads overlap)
  double pv[QS];
#pragma unroll
  for (int p2 = 0; p2 < QS / 2; ++p2) {
    const int4 c4 = __ldg(rec + 2 + p2);
    const uint32_t s0 = (static_cast<uint32_t>(c4.x) >> 23) & 127u, s1 = (static_cast<uint32_t>(c4.z) >> 23) & 127u;
    pv[2 * p2] = s0 == SLOT_NONE ? 0.0 : plane(s0);
    pv[2 * p2 + 1] = s1 == SLOT_NONE ? 0.0 : plane(s1);
  }
  double acc = 0.0;
  bool matched = false;
#pragma unroll
  for (int pp = 0; pp < QS; ++pp) {
    double v = pv[pp];
    if ((own_bits >> pp) & 1u) v = myval;
    if ((hits >> pp) & 1u) {
      const int idx = row_find(ea.drow_term + r0, row_n, static_cast<uint32_t>(__ldg(ea.qinfo + q * QS + pp)));
      if (idx >= 0) { matched = true; v = __ldg(ea.drow_val + r0 + idx); }
    }
    acc = __dadd_rn(acc, v);
  }
  if (ea.tp_count && matched) atomicAdd(ea.tp_count, 1);       // diagnostics
  member_emit(ea, q, acc, (static_cast<unsigned long long>(static_cast<uint32_t>(hd.w)) << 32) | static_cast<uint32_t>(hd.z), d);
}
// in place (a full queue): fp64 planes from global memory, pt = the doc's entry in plane 0 of its tile
template <int QS>
__device__ __noinline__ void member_exact(const ExactArgs ea, const double* __restrict__ pt, int64_t d, double myval,
                                          const int4* __restrict__ rec, uint32_t hits) {
  exact_pair<QS>(ea, [&](uint32_t slot) { return __ldg(pt + slot * PLANE_TILE); }, d, myval, rec, hits);
}

// The set-aside pairs of a member-pass launch, one per thread, thousands in flight per SM: chains of dependent loads
// that would idle a persistent CTA.  grid (x, member-pass CTAs); planes are read from global memory.
struct __align__(8) MjDefer { uint32_t doc, rec, hits, pad; double myval; };   // the pair's doc and own contribution travel with it
template <int QS>
__global__ void __launch_bounds__(256)
bm25_member_exact_kernel(const Bm25Dev ix, const BatchScratch sc) {
  constexpr int rs4 = 2 + QS / 2;
  const MjDefer* defer = sc.defer + static_cast<size_t>(blockIdx.y) * MJ_DEF;
  const int n = min(sc.def_cnt[blockIdx.y], MJ_DEF);
  const ExactArgs ea{ix.drow_off, ix.drow_term, ix.drow_val, sc.qinfo, sc.c_key, sc.c_doc, sc.c_cnt, sc.n_nonpos, sc.qs, ix.n_planes,
                     sc.stats_on ? sc.counters + 4 : nullptr};
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const MjDefer de = defer[k];
    const double* pt = ix.planes + static_cast<int64_t>(de.doc >> PLANE_TILE_LOG2) * ix.n_planes * PLANE_TILE + plane_pos(de.doc & (PLANE_TILE - 1));
    exact_pair<QS>(ea, [&](uint32_t slot) { return __ldg(pt + slot * PLANE_TILE); }, static_cast<int64_t>(de.doc), de.myval,
               sc.bt_rec + static_cast<size_t>(de.rec) * rs4, de.hits);
  }
}

// grid: persistent CTAs over the units of tiles [tile_lo, tile_hi); a unit is 1/sub of a tile's sparse postings.
// Per unit: stage the tile of all planes (bulk copy) and build the per-doc filters; then, `ch` postings at a time
// (term-major: the postings of a term inside the tile form a run), count the (record, doc) pairs per run
// (records of the term in the batch x docs of the run), scan, and give every thread an equal, contiguous range of
// pairs in run-major / record-major order: the thread decodes a record into registers once and applies it to every
// doc of the run.  The pair is SCREENED in fp32: the staged tile holds the planes rounded up to fp32, the record's own
// token comes from the posting (rounded up), every other sparse token adds +0.0, and the sum runs in query-token order
// with upward rounding -- a rigorous upper bound of the in-order fp64 score (fp64 rounding is monotone), at half the
// shared-memory bytes of an fp64 tile.  Only pairs whose bound exceeds the query's running bound -- a few hundred per
// query and round -- and pairs where the doc's filter says another sparse token of the query may be present (with the
// positions in question) are queued for bm25_member_exact_kernel, which computes the exact fp64 score.
// QS = tokens per query rounded up to 4 (<= 32).
template <int QS>
__global__ void __launch_bounds__(MJ_THREADS, 1)
bm25_member_kernel(const Bm25Dev ix, const BatchScratch sc, int64_t tile_lo, int64_t tile_hi, int sub, int ch, int dbg) {
  extern __shared__ __align__(16) uint8_t mj_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = ix.n_planes;
  float* pl = reinterpret_cast<float*>(mj_smem);                                   // [P + 1][MJ_PSTRIDE]
  uint32_t* bloom = reinterpret_cast<uint32_t*>(mj_smem + static_cast<size_t>(P + 1) * MJ_PSTRIDE * 4);   // [256][16]
  uint32_t* pscan = bloom + PLANE_TILE * MJ_BLOOM_W;                               // [ch + 4] exclusive scan of the runs' pair counts
  uint32_t* cbase = pscan + ch + 4;                                                // [ch] first record of a run's term (at its head)
  uint32_t* pinfo = cbase + ch;                                                    // [ch] doc in tile | run length << 8 | head << 31
  MjDefer* defer = sc.defer + static_cast<size_t>(blockIdx.x) * MJ_DEF;
  __shared__ __align__(8) uint64_t fill_bar;
  __shared__ uint32_t wsum[32];
  __shared__ int n_def_s;
  if (tid == 0) { ptx::mbar_init(&fill_bar, 1); ptx::fence_barrier_init(); n_def_s = 0; }
  for (int i = tid; i < MJ_PSTRIDE; i += MJ_THREADS) pl[static_cast<size_t>(P) * MJ_PSTRIDE + i] = 0.0f;
  constexpr uint32_t BLOOM_BYTES = PLANE_TILE * MJ_BLOOM_W * 4u;
  uint32_t fill_phase = 0;
  const int64_t units = (tile_hi - tile_lo) * sub;
  constexpr int RS4 = 2 + QS / 2;
  const ExactArgs ea{ix.drow_off, ix.drow_term, ix.drow_val, sc.qinfo, sc.c_key, sc.c_doc, sc.c_cnt, sc.n_nonpos, sc.qs, P, nullptr};
  for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
    const int64_t tile = tile_lo + u / sub;
    const int part = static_cast<int>(u % sub);
    __syncthreads();                                           // the previous unit is done; fill_bar is initialised
    if (warp == 0) {
      // the tile of all planes (one 2 KB row at a time into the padded rows) and the tile's per-doc filters: the copies
      // are issued by the 32 lanes side by side, all completing on one mbarrier
      ptx::fence_proxy_async_smem();                            // every issuing lane: earlier generic reads before its async writes
      if (lane == 0) {
        ptx::mbar_expect_tx(&fill_bar, static_cast<uint32_t>(P) * PLANE_TILE * 4u + BLOOM_BYTES);
        ptx::bulk_load_1d(bloom, ix.tt_bloom + tile * (PLANE_TILE * MJ_BLOOM_W), BLOOM_BYTES, &fill_bar);
      }
      __syncwarp();
      const char* src = reinterpret_cast<const char*>(ix.planes32 + tile * P * PLANE_TILE);
      for (int r = lane; r < P; r += 32)
        ptx::bulk_load_1d(mj_smem + static_cast<size_t>(r) * MJ_PSTRIDE * 4, src + static_cast<size_t>(r) * PLANE_TILE * 4,
                          PLANE_TILE * 4u, &fill_bar);
    }
    const uint32_t t_lo = __ldg(ix.tt_off + tile), t_hi = __ldg(ix.tt_off + tile + 1);
    ptx::mbar_wait(&fill_bar, fill_phase);
    fill_phase ^= 1u;
    __syncthreads();
    // this unit's slice of the tile's postings
    const uint32_t o_lo = t_lo + static_cast<uint32_t>((static_cast<uint64_t>(t_hi - t_lo) * part) / sub);
    const uint32_t o_hi = t_lo + static_cast<uint32_t>((static_cast<uint64_t>(t_hi - t_lo) * (part + 1)) / sub);
    const int64_t doc_base = tile * PLANE_TILE;
    for (uint32_t c0 = o_lo; c0 < o_hi; c0 += ch) {
      const int n_ch = static_cast<int>(min(static_cast<uint32_t>(ch), o_hi - c0));
      // ---- runs (thread t owns postings MJ_PPT * t ..): the index knows where a term's run starts and how far it goes
      // (tt_info); a run head -- and the first posting of the chunk, which may sit inside a run -- looks up the term's
      // records in the batch: pairs of the run = records x docs
      uint32_t cnt[MJ_PPT];
      uint32_t tsum = 0u;
#pragma unroll
      for (int e = 0; e < MJ_PPT; ++e) {
        const int i = tid * MJ_PPT + e;
        cnt[e] = 0u;
        if (i < n_ch) {
          const uint32_t info = __ldg(ix.tt_info + c0 + i);
          uint32_t out = info & 0xFFu;
          if ((info >> 31) || i == 0) {
            const uint2 be = __ldg(sc.bt_ent + __ldg(ix.tt_term + c0 + i));
            const uint32_t len = min(((info >> 8) & 0xFFu) + 1u, static_cast<uint32_t>(n_ch - i));
            cbase[i] = be.x;
            out |= (len << 8) | 0x80000000u;
            cnt[e] = be.y * len;
          }
          pinfo[i] = out;
        }
        tsum += cnt[e];
      }
      uint32_t incl = tsum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      if (lane == 31) wsum[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        const uint32_t w = lane < MJ_THREADS / 32 ? wsum[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
        wsum[lane] = wi - w;
        if (lane == 31) pscan[n_ch] = wi;
      }
      __syncthreads();
      {
        uint32_t run = wsum[warp] + incl - tsum;
#pragma unroll
        for (int e = 0; e < MJ_PPT; ++e) {
          const int i = tid * MJ_PPT + e;
          if (i < n_ch) pscan[i] = run;
          run += cnt[e];
        }
      }
      __syncthreads();
      // ---- pairs: thread t scores pairs [t * R, (t + 1) * R) of the chunk
      const uint32_t T = pscan[n_ch];
      const uint32_t R = (T + MJ_THREADS - 1) / MJ_THREADS;
      uint32_t j = tid * R;
      const uint32_t j_end = min(T, j + R);
      if (j < j_end && !(dbg & 4)) {
        int i0 = seg_of(pscan, n_ch, j);                       // head of the run holding pair j
        int len = static_cast<int>((pinfo[i0] >> 8) & 0x7FFFFFu);
        uint32_t n_rec = (pscan[i0 + 1] - pscan[i0]) / static_cast<uint32_t>(len);
        uint32_t k = (j - pscan[i0]) / static_cast<uint32_t>(len);                 // record of the run
        int di = static_cast<int>(j - pscan[i0] - k * static_cast<uint32_t>(len));  // doc of the run
        bool fresh = true;
        // the decoded record: per token the byte offset of its plane row (zero row when it is no plane term), the byte
        // offset of its filter word and the mask it must find there
        uint32_t po[QS], bw[QS], bm[QS];
        uint32_t own_bit = 0u, plane_bits = 0u, sparse_bits = 0u, r = 0u;
        int64_t q = 0;
        double bound_d = 0.0;
        for (; j < j_end; ++j) {
          if (fresh) {
            fresh = false;
            r = cbase[i0] + k;
            const int4* rec = sc.bt_rec + static_cast<size_t>(r) * RS4;
            const int4 hd = __ldg(rec);
            ptx::prefetch_l1(rec + RS4);                       // the next record of the run (records of a term are contiguous)
            plane_bits = static_cast<uint32_t>(hd.y);
            const int2 h1 = __ldg(reinterpret_cast<const int2*>(rec + 1));
            sparse_bits = static_cast<uint32_t>(h1.x);
            own_bit = static_cast<uint32_t>(h1.y);
#pragma unroll
            for (int p2 = 0; p2 < QS / 2; ++p2) {
              const int4 c4 = __ldg(rec + 2 + p2);
              po[2 * p2] = static_cast<uint32_t>(c4.x) & 0x1FFFFu;
              bw[2 * p2] = (static_cast<uint32_t>(c4.x) >> 17) & 63u;
              bm[2 * p2] = static_cast<uint32_t>(c4.y);
              po[2 * p2 + 1] = static_cast<uint32_t>(c4.z) & 0x1FFFFu;
              bw[2 * p2 + 1] = (static_cast<uint32_t>(c4.z) >> 17) & 63u;
              bm[2 * p2 + 1] = static_cast<uint32_t>(c4.w);
            }
            q = static_cast<uint32_t>(hd.x) >> 5;
            bound_d = key_f64((static_cast<unsigned long long>(static_cast<uint32_t>(hd.w)) << 32) | static_cast<uint32_t>(hd.z));
          }
          const int i = i0 + di;
          const int dl = pinfo[i] & 0xFF;
          const double myval = __ldg(ix.tt_val + c0 + i);
          const char* pl_doc = reinterpret_cast<const char*>(pl + pos32(static_cast<uint32_t>(dl)));
          const char* brow = reinterpret_cast<const char*>(bloom + dl * MJ_BLOOM_W);
          const float myub = __double2float_ru(myval);
          float ub = 0.0f;
          uint32_t hits = 0u;
#pragma unroll
          for (int p = 0; p < QS; ++p) {
            // only real plane / sparse tokens touch shared memory: random addresses across a warp, fewer lanes = fewer bank conflicts
            if (plane_bits & (1u << p)) ub = __fadd_ru(ub, *reinterpret_cast<const float*>(pl_doc + po[p]));
            if (own_bit & (1u << p)) ub = __fadd_ru(ub, myub);
            if ((sparse_bits & (1u << p)) && (*reinterpret_cast<const uint32_t*>(brow + bw[p]) & bm[p]) == bm[p]) hits |= 1u << p;
          }
          // ub >= the exact score when no other sparse token matches.  A doc that may match another one (hits) or whose
          // bound beats the query's running bound gets its exact fp64 score from the queue's kernel; a bound that is not
          // positive settles the pair as a member that can never be a candidate.
          const bool maybe = hits != 0u || static_cast<double>(ub) > bound_d;
          if (maybe && !(dbg & 1)) {
            const int at = atomicAdd(&n_def_s, 1);
            if (at < MJ_DEF) { defer[at] = MjDefer{static_cast<uint32_t>(doc_base + dl), r, hits, 0u, myval}; }
            else member_exact<QS>(ea, ix.planes + tile * P * PLANE_TILE + plane_pos(static_cast<uint32_t>(dl)), doc_base + dl, myval,
                              sc.bt_rec + static_cast<size_t>(r) * RS4, hits);
          } else if (!maybe && !(ub > 0.0f)) {
            atomicAdd(sc.n_nonpos + q, 1);
          }
          if (++di == len) {                                   // next record of the run, or the next run that has records
            di = 0;
            fresh = true;
            if (++k == n_rec && j + 1 < j_end) {
              k = 0;
              do { i0 += len; len = static_cast<int>((pinfo[i0] >> 8) & 0x7FFFFFu); } while (pscan[i0 + 1] == pscan[i0]);
              n_rec = (pscan[i0 + 1] - pscan[i0]) / static_cast<uint32_t>(len);
            }
          }
        }
      }
      __syncthreads();                                         // pscan / cbase / pinfo are rewritten by the next chunk
    }
  }
  __syncthreads();
  if (tid == 0) { sc.def_cnt[blockIdx.x] = n_def_s; atomicAdd(&sc.counters[0], n_def_s); }   // this CTA's queue: bm25_member_exact_kernel finishes it
}

// Between rounds: cut a query's candidate list back to its k best and publish the k-th key as the next bound.
// After the last round: list A (sorted), the screening threshold and the query's state.  One CTA per query.
// The k best are SELECTED, not sorted out of the whole list: a 4096-bin histogram of the scores (128 bins per octave)
// finds the bin holding the k-th score; everything above it is in, and only that bin's entries are sorted.
__device__ __forceinline__ int key_bin(unsigned long long key) {          // same bins as h3_bin()
  const long long t = static_cast<long long>(key >> 45) - static_cast<long long>(H3_BASE);
  return static_cast<int>(t < 0 ? 0 : (t > H3_BINS - 1 ? H3_BINS - 1 : t));
}
constexpr int CP_SLOTS = C_CAP + 256;
constexpr int CP_SMEM = CP_SLOTS * 12 + H3_BINS * 4;
__global__ void __launch_bounds__(256)
bm25_compact_kernel(int64_t nq, int k, int last, const BatchScratch sc) {
  extern __shared__ __align__(16) uint8_t cp_smem[];
  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int raw = sc.c_cnt[q];
  const bool ovf = sc.ovf[q] != 0 || raw > C_CAP;
  auto publish = [&](unsigned long long bound) {              // the next round reads the bound from the query's records
    const int rs4 = rec_int4s(sc.qs);
    if (tid < sc.qs) {
      const uint32_t r = sc.rec_of[q * sc.qs + tid];
      if (r != 0xFFFFFFFFu)
        *reinterpret_cast<unsigned long long*>(reinterpret_cast<int*>(sc.bt_rec + static_cast<size_t>(r) * rs4) + 2) = bound;
    }
  };
  if (!last) {
    if (ovf) {
      if (tid == 0) { sc.ovf[q] = 1; sc.gthr[q] = ~0ull; sc.c_cnt[q] = 0; }
      publish(~0ull);
      return;
    }
    if (raw <= k) return;
  }
  const int n = ovf ? 0 : raw;
  unsigned long long* key = reinterpret_cast<unsigned long long*>(cp_smem);
  uint32_t* doc = reinterpret_cast<uint32_t*>(cp_smem + static_cast<size_t>(CP_SLOTS) * 8);
  uint32_t* hist = doc + CP_SLOTS;
  __shared__ int s_bin, s_hi, s_nhi, s_nb;
  __shared__ uint32_t s_w[8];
  int cnt = min(n, k);
  if (n > k) {
    for (int i = tid; i < H3_BINS; i += 256) hist[i] = 0u;
    if (tid == 0) { s_nhi = 0; s_nb = 0; }
    __syncthreads();
    for (int i = tid; i < n; i += 256) atomicAdd(&hist[key_bin(sc.c_key[q * C_CAP + i])], 1u);
    __syncthreads();
    // thread t owns the 16 bins [4096 - 16 (t + 1), 4096 - 16 t): counts from the top
    uint32_t mine = 0u;
    const int b_top = H3_BINS - 16 * tid;
#pragma unroll
    for (int j = 1; j <= 16; ++j) mine += hist[b_top - j];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t above = incl - mine;
    for (int w = 0; w < warp; ++w) above += s_w[w];
    if (above < static_cast<uint32_t>(k) && above + mine >= static_cast<uint32_t>(k)) {   // the k-th score lies in one of my bins
      int b = b_top;
      uint32_t c = above;
      do { --b; c += hist[b]; } while (c < static_cast<uint32_t>(k));
      s_bin = b; s_hi = static_cast<int>(c - hist[b]);
    }
    __syncthreads();
    const int B = s_bin, c_hi = s_hi;                            // c_hi < k entries lie in bins above B
    for (int i = tid; i < n; i += 256) {
      const unsigned long long kk = sc.c_key[q * C_CAP + i];
      const int b = key_bin(kk);
      if (b > B) { const int at = atomicAdd(&s_nhi, 1); key[at] = kk; doc[at] = sc.c_doc[q * C_CAP + i]; }
      else if (b == B) { const int at = c_hi + atomicAdd(&s_nb, 1); key[at] = kk; doc[at] = sc.c_doc[q * C_CAP + i]; }
    }
    __syncthreads();
    const int nb = s_nb;
    if (nb > k - c_hi) {                                       // more than needed in bin B: the best of them by (key desc, doc asc)
      int n2 = 2;
      while (n2 < nb) n2 <<= 1;
      for (int i = nb + tid; i < n2; i += 256) { key[c_hi + i] = 0ull; doc[c_hi + i] = 0xFFFFFFFFu; }
      cta_sort_desc(key + c_hi, doc + c_hi, n2);
    }
  } else {
    for (int i = tid; i < n; i += 256) { key[i] = sc.c_key[q * C_CAP + i]; doc[i] = sc.c_doc[q * C_CAP + i]; }
    __syncthreads();
  }
  // key/doc[0, cnt) now hold the cnt best entries (unordered above bin B)
  if (!last) {
    // the k-th key = the smallest selected one
    unsigned long long kmin = ~0ull;
    for (int i = tid; i < cnt; i += 256) kmin = min(kmin, key[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    __shared__ unsigned long long s_min[8];
    if (lane == 0) s_min[warp] = kmin;
    __syncthreads();
    kmin = s_min[0];
    for (int w = 1; w < 8; ++w) kmin = min(kmin, s_min[w]);
    for (int i = tid; i < cnt; i += 256) { sc.c_key[q * C_CAP + i] = key[i]; sc.c_doc[q * C_CAP + i] = doc[i]; }
    if (tid == 0) { sc.c_cnt[q] = cnt; if (cnt >= k) sc.gthr[q] = kmin; }
    if (cnt >= k) publish(kmin);
    return;
  }
  int n2 = 2;
  while (n2 < cnt) n2 <<= 1;
  __syncthreads();
  for (int i = cnt + tid; i < n2; i += 256) { key[i] = 0ull; doc[i] = 0xFFFFFFFFu; }
  cta_sort_desc(key, doc, n2);
  for (int i = tid; i < k; i += blockDim.x) {
    sc.a_key[q * k + i] = i < cnt ? key[i] : 0ull;
    sc.a_doc[q * k + i] = i < cnt ? doc[i] : 0xFFFFFFFFu;
  }
  if (tid == 0) {
    const bool has_planes = sc.qd[q * sc.qd_stride] != 0xFF;
    sc.a_cnt[q] = cnt;
    int h3 = H3_MAX;
    if (!has_planes && !ovf)    { sc.state[q] = 1; sc.thr[q] = THR_SKIP; }
    else if (cnt >= k && !ovf)  { sc.state[q] = 0; sc.thr[q] = key[k - 1]; }
    else if (!ovf && (h3 = atomicAdd(&sc.counters[2], 1)) < H3_MAX) {
      sc.state[q] = 3; sc.thr[q] = THR_SKIP; sc.h3_list[h3] = static_cast<int>(q);
    } else {
      sc.state[q] = 2; sc.thr[q] = THR_SKIP;
      sc.flag_list[atomicAdd(&sc.counters[1], 1)] = static_cast<int>(q);
    }
  }
}

// ----------------------------------------------------------------------------
// 2. screening pass: grid (tile sets, query groups), one CTA per SM, two fp32 plane tiles in flight.
// ----------------------------------------------------------------------------
// append the docs of one tile whose in-order fp64 plane score reaches thr_d and that are not members of q
struct RecheckArgs {
  const double* planes; const uint32_t* drow_off; const uint32_t* drow_term; const int32_t* qinfo; const uint8_t* qd;
  int* e_cnt; unsigned long long* e_key; uint32_t* e_doc; int P, qs, qd_stride;
};
__device__ __noinline__ void screen_recheck(const RecheckArgs ra, int64_t q, int64_t tile, double thr_d, int lane) {
  const int P = ra.P;
  double acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.0;
  for (int i = 0; i < ra.qd_stride; ++i) {
    const uint32_t ps = ra.qd[q * ra.qd_stride + i];
    if (ps == 0xFFu) break;
    const double2* row = reinterpret_cast<const double2*>(ra.planes + (tile * P + ps) * PLANE_TILE) + lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double2 v = __ldg(row + j * 32);
      acc[2 * j] = __dadd_rn(acc[2 * j], v.x);
      acc[2 * j + 1] = __dadd_rn(acc[2 * j + 1], v.y);
    }
  }
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int64_t d = tile * PLANE_TILE + 8 * lane + r;
    bool p = acc[r] >= thr_d && acc[r] > 0.0;
    if (p) p = !is_member(ra.drow_off, ra.drow_term, ra.qinfo, ra.qs, q, d);
    const uint32_t m = __ballot_sync(0xffffffffu, p);
    if (m == 0u) continue;
    int base = 0;
    if (lane == 0) base = atomicAdd(&ra.e_cnt[q], __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (p) {
      const int at = base + __popc(m & lt);
      if (at < E_CAP) { ra.e_key[q * E_CAP + at] = f64_key(acc[r]); ra.e_doc[q * E_CAP + at] = static_cast<uint32_t>(d); }
    }
  }
}

__global__ void __launch_bounds__(SC_WARPS * 32, 1)
bm25_screen_kernel(const Bm25Dev ix, int64_t nq, int64_t q_per_group, const BatchScratch sc, int grab) {
  extern __shared__ __align__(16) uint8_t sn_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int P = ix.n_planes, Pp = (P + 1) & ~1;
  const uint32_t tile_bytes = static_cast<uint32_t>(P) * PLANE_TILE * 4u;
  double* tmax_s = reinterpret_cast<double*>(sn_smem + 2 * static_cast<size_t>(tile_bytes));       // [2][Pp]
  ulonglong2* qrec = reinterpret_cast<ulonglong2*>(tmax_s + 2 * Pp);                               // [q_per_group]
  int* qidx = reinterpret_cast<int*>(qrec + q_per_group);                                          // [q_per_group]
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ int n_active;
  __shared__ int next_a[2];                                    // next active query of the tile in buffer b (warps take turns)
  const int64_t q0 = blockIdx.y * q_per_group;
  const int64_t q1 = min(nq, q0 + q_per_group);
  const int qw = sc.qd_stride >> 3;                            // 8-byte words of packed plane slots per query
  const unsigned long long* qd64 = reinterpret_cast<const unsigned long long*>(sc.qd);
  if (threadIdx.x == 0) { n_active = 0; ptx::mbar_init(&bar[0], 1); ptx::mbar_init(&bar[1], 1); ptx::fence_barrier_init(); }
  __syncthreads();
  // Upper bounds: fp64 addition is monotone, so adding the plane maxima in the query's token order bounds the
  // in-order score of every doc that matches plane tokens only.  A query whose bound over the whole corpus is
  // below its threshold needs no screening at all; the per-tile bound prunes (query, tile) pairs of the rest.
  for (int64_t q = q0 + threadIdx.x; q < q1; q += SC_WARPS * 32) {
    const unsigned long long t = __ldg(sc.thr + q);
    if (sc.state[q] != 0 || t == THR_SKIP) continue;
    const double thr_d = key_f64(t);
    double ub = 0.0;
    for (int i = 0; i < sc.qd_stride; ++i) {
      const uint32_t ps = sc.qd[q * sc.qd_stride + i];
      if (ps == 0xFFu) break;
      ub = __dadd_rn(ub, __ldg(ix.plane_gmax + ps));
    }
    if (ub >= thr_d) {
      const int a = atomicAdd(&n_active, 1);
      qrec[a] = make_ulonglong2(static_cast<unsigned long long>(__double_as_longlong(thr_d)), __ldg(qd64 + q * qw));
      qidx[a] = static_cast<int>(q);
    }
  }
  __syncthreads();
  const int na = n_active;
  if (na == 0) return;
  auto issue = [&](int64_t tile, int b) {                      // thread 0: hand one fp32 tile to the bulk-copy engine
    const char* src = reinterpret_cast<const char*>(ix.planes32 + tile * P * PLANE_TILE);
    uint8_t* dst = sn_smem + static_cast<size_t>(b) * tile_bytes;
    ptx::fence_proxy_async_smem();
    ptx::mbar_expect_tx(&bar[b], tile_bytes);
    for (uint32_t off = 0; off < tile_bytes; off += 32768u)
      ptx::bulk_load_1d(dst + off, src + off, min(32768u, tile_bytes - off), &bar[b]);
  };
  int64_t tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < ix.n_tiles) issue(tile, 0);
  uint32_t phases = 0u;
  for (int b = 0; tile < ix.n_tiles; tile += gridDim.x, b ^= 1) {
    // the other buffer was last read in the previous iteration, which ended with a CTA barrier
    if (threadIdx.x == 0 && tile + gridDim.x < ix.n_tiles) issue(tile + gridDim.x, b ^ 1);
    if (threadIdx.x < P) tmax_s[b * Pp + threadIdx.x] = __ldg(ix.plane_tmax + tile * P + threadIdx.x);
    if (threadIdx.x == 0) next_a[b] = 0;
    ptx::mbar_wait(&bar[b], (phases >> b) & 1u);
    phases ^= 1u << b;
    __syncthreads();
    const float* T = reinterpret_cast<const float*>(sn_smem + static_cast<size_t>(b) * tile_bytes);
    const double* tm = tmax_s + b * Pp;
    while (true) {                                             // warps take SC_GRAB active queries at a time
      int a0 = 0;
      if (lane == 0) a0 = atomicAdd(&next_a[b], grab);
      a0 = __shfl_sync(0xffffffffu, a0, 0);
      if (a0 >= na) break;
      // this tile's bound for SC_GRAB queries at once, one query per lane (same token order as the scores)
      const int a = a0 + lane;
      ulonglong2 rec = make_ulonglong2(0ull, ~0ull);
      int q_mine = 0;
      bool alive = false;
      if (lane < grab && a < na) {
        rec = qrec[a];                                         // {k-th score of A as fp64 bits, packed plane slots}
        q_mine = qidx[a];
        double ub = tm[static_cast<uint32_t>(rec.y) & 0xFFu];
        unsigned long long pb = rec.y;
        for (int i = 1; i < sc.qd_stride; ++i) {
          if ((i & 7) == 0) pb = __ldg(qd64 + static_cast<int64_t>(q_mine) * qw + (i >> 3)); else pb >>= 8;
          const uint32_t s2 = static_cast<uint32_t>(pb) & 0xFFu;
          if (s2 == 0xFFu) break;
          ub = __dadd_rn(ub, tm[s2]);
        }
        alive = ub >= __longlong_as_double(static_cast<long long>(rec.x));
      }
      uint32_t todo = __ballot_sync(0xffffffffu, alive);
      while (todo != 0u) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1u;
        const double thr_d = __longlong_as_double(static_cast<long long>(__shfl_sync(0xffffffffu, rec.x, src)));
        unsigned long long pk = __shfl_sync(0xffffffffu, rec.y, src);
        const int64_t q = __shfl_sync(0xffffffffu, q_mine, src);
        uint32_t ps = static_cast<uint32_t>(pk) & 0xFFu;
        // fp32 sums of rounded-up values with upward rounding: acc[r] >= the in-order fp64 score of doc 8*lane + r
        float acc[8];
        {
          const float4* row = reinterpret_cast<const float4*>(T + ps * PLANE_TILE) + lane;
          const float4 v0 = row[0], v1 = row[32];
          acc[0] = v0.x; acc[1] = v0.y; acc[2] = v0.z; acc[3] = v0.w;
          acc[4] = v1.x; acc[5] = v1.y; acc[6] = v1.z; acc[7] = v1.w;
        }
        for (int i = 1; i < sc.qd_stride; ++i) {
          if ((i & 7) == 0) pk = __ldg(qd64 + q * qw + (i >> 3)); else pk >>= 8;
          ps = static_cast<uint32_t>(pk) & 0xFFu;
          if (ps == 0xFFu) break;
          const float4* row = reinterpret_cast<const float4*>(T + ps * PLANE_TILE) + lane;
          const float4 v0 = row[0], v1 = row[32];
          acc[0] = __fadd_ru(acc[0], v0.x); acc[1] = __fadd_ru(acc[1], v0.y);
          acc[2] = __fadd_ru(acc[2], v0.z); acc[3] = __fadd_ru(acc[3], v0.w);
          acc[4] = __fadd_ru(acc[4], v1.x); acc[5] = __fadd_ru(acc[5], v1.y);
          acc[6] = __fadd_ru(acc[6], v1.z); acc[7] = __fadd_ru(acc[7], v1.w);
        }
        const float thr_f = __double2float_rd(thr_d);
        bool pass = false;
#pragma unroll
        for (int r = 0; r < 8; ++r) pass |= acc[r] >= thr_f;
        if (__ballot_sync(0xffffffffu, pass) == 0u) continue;
        // some bound reaches the threshold: the (query, tile) pair is recomputed in fp64 by bm25_recheck_kernel (its
        // chains of dependent global loads would stall this CTA's tile barrier; queued, thousands run side by side)
        int at = 0;
        if (lane == 0) at = atomicAdd(&sc.counters[3], 1);
        at = __shfl_sync(0xffffffffu, at, 0);
        if (at < RQ_CAP) { if (lane == 0) sc.rq[at] = make_uint2(static_cast<uint32_t>(q), static_cast<uint32_t>(tile)); }
        else screen_recheck(RecheckArgs{ix.planes, ix.drow_off, ix.drow_term, sc.qinfo, sc.qd, sc.e_cnt, sc.e_key, sc.e_doc, P, sc.qs,
                                        sc.qd_stride}, q, tile, thr_d, lane);
      }
    }
    __syncthreads();
  }
}

// one warp per queued (query, tile) pair
__global__ void __launch_bounds__(256)
bm25_recheck_kernel(const Bm25Dev ix, const BatchScratch sc) {
  const int lane = threadIdx.x & 31;
  const int n = min(sc.counters[3], RQ_CAP);
  const int wpg = (blockDim.x >> 5) * gridDim.x;
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += wpg) {
    const uint2 e = sc.rq[i];
    screen_recheck(RecheckArgs{ix.planes, ix.drow_off, ix.drow_term, sc.qinfo, sc.qd, sc.e_cnt, sc.e_key, sc.e_doc, ix.n_planes, sc.qs,
                               sc.qd_stride}, e.x, e.y, key_f64(sc.thr[e.x]), lane);
  }
}

// The per-query fallback has a fixed latency of ~0.2 ms per batch however few queries take it.  A large batch that lists
// only a few queries for the histogram path already pays that latency for its overflowed queries: hand the listed ones
// over as well (their low thresholds would add more screening work than the fallback costs).  One block.
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
// 2a. histogram path: queries whose list A holds fewer than k docs although they have plane tokens (frequent on
// small shards and tiny vocabularies).  EMIT = false: histogram of the plane-only scores of the docs of every
// tile_step-th tile (a member's plane-only score is counted too: the threshold kernel asks for correspondingly more
// docs); the threshold derived from it is what the screening pass then uses for the query.  (EMIT = true, a second
// sweep that emits the docs at or above the threshold itself, is no longer launched.)  Returns at once when the list is empty.
// ----------------------------------------------------------------------------
template <bool EMIT>
__global__ void __launch_bounds__(PL_WARPS * 32, 1)
bm25_plane_h3_kernel(const Bm25Dev ix, const BatchScratch sc, int tile_step) {
  extern __shared__ __align__(16) double pll_smem[];           // [n_planes][256]
  const int n3 = min(sc.counters[2], H3_MAX);
  if (n3 == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int P = ix.n_planes;
  const uint32_t lt = (1u << lane) - 1u;
  if (!EMIT && n3 <= H3_DIRECT) {
    // a handful of queries: staging whole tiles (96 planes) for them costs more than it saves -- one warp per
    // (query, sampled tile) reads just the query's plane rows from global memory (2 KB each, coalesced)
    const int64_t n_samp = (ix.n_tiles + tile_step - 1) / tile_step;
    const int64_t tasks = n_samp * n3;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * PL_WARPS + warp; t < tasks; t += static_cast<int64_t>(gridDim.x) * PL_WARPS) {
      const int idx = static_cast<int>(t % n3);
      const int64_t tile = (t / n3) * tile_step;
      const int64_t q = sc.h3_list[idx];
      double acc[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] = 0.0;
      for (int i = 0; i < sc.qd_stride; ++i) {
        const uint32_t ps = sc.qd[q * sc.qd_stride + i];
        if (ps == 0xFFu) break;
        const double2* row = reinterpret_cast<const double2*>(ix.planes + (tile * P + ps) * PLANE_TILE) + lane;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double2 v = __ldg(row + j * 32);
          acc[2 * j] = __dadd_rn(acc[2 * j], v.x);
          acc[2 * j + 1] = __dadd_rn(acc[2 * j + 1], v.y);
        }
      }
      uint32_t* H = sc.hist + static_cast<size_t>(idx) * H3_BINS;
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (acc[r] > 0.0) atomicAdd(&H[h3_bin(acc[r])], 1u);
    }
    return;
  }
  for (int64_t tile = static_cast<int64_t>(blockIdx.x) * tile_step; tile < ix.n_tiles; tile += static_cast<int64_t>(gridDim.x) * tile_step) {
    __syncthreads();
    {
      const double2* src = reinterpret_cast<const double2*>(ix.planes + tile * P * PLANE_TILE);
      double2* dst = reinterpret_cast<double2*>(pll_smem);
      for (int i = threadIdx.x; i < P * (PLANE_TILE / 2); i += PL_WARPS * 32) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    for (int idx = warp; idx < n3; idx += PL_WARPS) {
      const int64_t q = sc.h3_list[idx];
      double thr_d = 0.0;
      if (EMIT) {
        const unsigned long long t = sc.thr[q];
        if (t == THR_SKIP) continue;
        thr_d = key_f64(t);
      }
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
      if constexpr (!EMIT) {
        uint32_t* H = sc.hist + static_cast<size_t>(idx) * H3_BINS;
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (acc[r] > 0.0) atomicAdd(&H[h3_bin(acc[r])], 1u);
      } else {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int64_t d = tile * PLANE_TILE + 8 * lane + r;
        bool p = acc[r] > 0.0 && acc[r] >= thr_d;
        if (p) p = !is_member(ix.drow_off, ix.drow_term, sc.qinfo, sc.qs, q, d);
        const uint32_t m = __ballot_sync(0xffffffffu, p);
        if (m == 0u) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(&sc.e_cnt[q], __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (p) {
          const int at = base + __popc(m & lt);
          if (at < E_CAP) { sc.e_key[q * E_CAP + at] = f64_key(acc[r]); sc.e_doc[q * E_CAP + at] = static_cast<uint32_t>(d); }
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
  // the histogram counts every doc by its plane-only score, members included: asking for k more than the number of
  // members leaves at least k non-members at or above the edge
  const int need = k + sc.a_cnt[q] + sc.n_nonpos[q];
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

static thread_local int64_t g_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // queries, fallback queries, emissions, max emissions, rechecks, set-aside pairs, of which matched, -

static int batch_set_attrs() {
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    err = cudaFuncSetAttribute(bm25_member_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
#define ARB_MJ(QS_) if (err == cudaSuccess) err = cudaFuncSetAttribute(bm25_member_kernel<QS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    ARB_MJ(8) ARB_MJ(12) ARB_MJ(16) ARB_MJ(20) ARB_MJ(24) ARB_MJ(28) ARB_MJ(32)
#undef ARB_MJ
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(bm25_screen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(bm25_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CP_SMEM);
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(bm25_plane_h3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_PLANES * PLANE_TILE * 8);
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(bm25_plane_h3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_PLANES * PLANE_TILE * 8);
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(bm25_flag_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16);
  });
  ARB_CUDA(err);
  return bm25_set_smem_attrs();
}

static int batch_chunk(const Bm25Dev& ix, const int32_t* d_terms, int64_t nq, int mq, int k, int64_t id_base,
                       double* d_scores, int64_t* d_ids, cudaStream_t st) {
  const int sms = num_sms();
  const bool planes = ix.n_planes > 0;
  BatchScratch sc{};
  sc.nq = nq;
  sc.qs = (mq + 3) & ~3;
  sc.qd_stride = (mq + 7) & ~7;
  AsyncBuf b_defer, b_rq;
  ARB_CUDA(b_defer.alloc(static_cast<size_t>(sms) * MJ_DEF * sizeof(MjDefer) + static_cast<size_t>(sms) * 4, st));
  sc.defer = b_defer.as<MjDefer>();
  sc.def_cnt = reinterpret_cast<int*>(sc.defer + static_cast<size_t>(sms) * MJ_DEF);
  AsyncBuf b_qdup, b_qinfo, b_bt, b_pair, b_ckey, b_cdoc, b_akey, b_adoc, b_i32, b_u64, b_u8, b_ekey, b_edoc, b_hist;
  const size_t nt = static_cast<size_t>(std::max<int64_t>(ix.n_terms, 1));
  ARB_CUDA(b_qinfo.alloc(static_cast<size_t>(nq) * sc.qs * 4, st));
  ARB_CUDA(b_bt.alloc(3 * nt * 4, st));                                             // bt_ent | bt_cur
  ARB_CUDA(b_pair.alloc(static_cast<size_t>(nq) * sc.qs * (32 + sc.qs * 8) + static_cast<size_t>(nq) * sc.qs * 4, st));   // bt_rec | rec_of
  ARB_CUDA(b_ckey.alloc(static_cast<size_t>(nq) * C_CAP * 8, st));
  ARB_CUDA(b_cdoc.alloc(static_cast<size_t>(nq) * C_CAP * 4, st));
  ARB_CUDA(b_akey.alloc(static_cast<size_t>(nq) * k * 8, st));
  ARB_CUDA(b_adoc.alloc(static_cast<size_t>(nq) * k * 4, st));
  ARB_CUDA(b_i32.alloc((static_cast<size_t>(nq) * 5 + 8 + H3_MAX) * 4, st));        // c_cnt | n_nonpos | e_cnt | counters | a_cnt | flag_list | h3_list
  ARB_CUDA(b_u64.alloc(static_cast<size_t>(nq) * 2 * 8, st));                       // gthr | thr
  ARB_CUDA(b_u8.alloc(static_cast<size_t>(nq) * (sc.qd_stride + 2) + 16, st));      // qd | state | ovf
  ARB_CUDA(b_qdup.alloc(static_cast<size_t>(nq) * 4, st));
  sc.qdup = b_qdup.as<uint32_t>();
  sc.qinfo = b_qinfo.as<int32_t>();
  sc.bt_ent = b_bt.as<uint2>();
  sc.bt_cur = b_bt.as<uint32_t>() + 2 * nt;
  sc.bt_rec = b_pair.as<int4>();
  sc.rec_of = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(b_pair.p) + static_cast<size_t>(nq) * sc.qs * (32 + sc.qs * 8));
  sc.c_key = b_ckey.as<unsigned long long>();
  sc.c_doc = b_cdoc.as<uint32_t>();
  sc.a_key = b_akey.as<unsigned long long>();
  sc.a_doc = b_adoc.as<uint32_t>();
  sc.c_cnt = b_i32.as<int>();
  sc.n_nonpos = sc.c_cnt + nq;
  sc.e_cnt = sc.n_nonpos + nq;
  sc.counters = sc.e_cnt + nq;
  sc.a_cnt = sc.counters + 8;
  sc.stats_on = std::getenv("ARB_BM25_STATS") ? 1 : 0;
  sc.flag_list = sc.a_cnt + nq;
  sc.h3_list = sc.flag_list + nq;
  sc.gthr = b_u64.as<unsigned long long>();
  sc.thr = sc.gthr + nq;
  sc.qd = b_u8.as<uint8_t>();                                   // 8-byte aligned rows first
  sc.state = sc.qd + static_cast<size_t>(nq) * sc.qd_stride;
  sc.ovf = sc.state + nq;
  if (planes) {
    ARB_CUDA(b_ekey.alloc(static_cast<size_t>(nq) * E_CAP * 8, st));
    ARB_CUDA(b_edoc.alloc(static_cast<size_t>(nq) * E_CAP * 4, st));
    ARB_CUDA(b_hist.alloc(static_cast<size_t>(H3_MAX) * H3_BINS * 4, st));
    ARB_CUDA(b_rq.alloc(static_cast<size_t>(RQ_CAP) * 8, st));
    sc.rq = b_rq.as<uint2>();
    sc.e_key = b_ekey.as<unsigned long long>();
    sc.e_doc = b_edoc.as<uint32_t>();
    sc.hist = b_hist.as<uint32_t>();
  }
  ARB_CUDA(cudaMemsetAsync(sc.bt_ent, 0, 3 * nt * 4, st));
  ARB_CUDA(cudaMemsetAsync(sc.c_cnt, 0, (static_cast<size_t>(nq) * 3 + 8) * 4, st));   // c_cnt, n_nonpos, e_cnt, counters
  ARB_CUDA(cudaMemsetAsync(sc.ovf, 0, static_cast<size_t>(nq), st));

  {
    ProfileScope ps(PROF_BM25_SPARSE, st);
    // ---- 0. query-term index
    bm25_qclassify_kernel<<<static_cast<unsigned>((nq + 127) / 128), 128, 0, st>>>(ix, d_terms, mq, nq, sc);
    ARB_LAUNCH_OK();
    bm25_bt_alloc_kernel<<<static_cast<unsigned>((nq * sc.qs + 255) / 256), 256, 0, st>>>(nq * sc.qs, sc);
    ARB_LAUNCH_OK();
    bm25_bt_fill_kernel<<<static_cast<unsigned>((nq * sc.qs + 255) / 256), 256, 0, st>>>(nq * sc.qs, sc, ix.n_planes);
    ARB_LAUNCH_OK();
    // ---- 1. member pass in rounds of growing size, a compaction after each.  The first round runs without a bound, so
    // every member with a positive score is a candidate: it covers ~64 tiles (16 K docs) whatever the corpus size, which
    // keeps even heavy queries below C_CAP; each later round covers 4x the docs before it, so with the bound of the rounds
    // before it a query collects ~3k new candidates per round, independent of the corpus size (an extra round at half the
    // corpus measured slower: its launches and compaction cost more than the ~100 candidates per query it saves).
    std::vector<int64_t> cuts;                                   // ascending round ends, in tiles
    if (ix.n_tiles <= 8) cuts = {ix.n_tiles};
    else if (ix.n_tiles < 1024) cuts = {(ix.n_tiles * 4 + 63) / 64, (ix.n_tiles * 16 + 63) / 64, ix.n_tiles};   // shards of up to 256 K docs: a round costs more than it prunes
    else {
      cuts = {ix.n_tiles};
      for (int64_t c = (ix.n_tiles + 3) / 4; ; c = (c + 3) / 4) { cuts.push_back(c); if (c <= 64) break; }
      std::reverse(cuts.begin(), cuts.end());
    }
    if (const char* v = std::getenv("ARB_BM25_ROUNDS")) { if (atoi(v) == 1) cuts = {ix.n_tiles}; }
    const int mj_ch = static_cast<int>(std::min<size_t>(MJ_CH_MAX, ((SMEM_MAX - mj_fixed_bytes(ix.n_planes)) / 12 - 4) & ~size_t(3)));
    const size_t mj_smem = mj_fixed_bytes(ix.n_planes) + static_cast<size_t>(mj_ch + 4) * 12;
    int64_t prev = 0;
    const int dbg = std::getenv("ARB_BM25_DEBUG") ? atoi(std::getenv("ARB_BM25_DEBUG")) : 0;   // timing experiments only (wrong results)
    for (size_t r = 0; r < cuts.size(); ++r) {
      const int64_t hi = std::max(prev, cuts[r]);
      const bool last = r + 1 == cuts.size();
      // a launch covers at most MJ_LAUNCH_TILES tiles: its queue of pairs for the exact kernel (MJ_DEF per CTA) then
      // holds what a cfg3-sized round produces, whatever the corpus size
      for (int64_t lo = prev; lo < hi; lo += MJ_LAUNCH_TILES) {
        const int64_t lhi = std::min<int64_t>(hi, lo + MJ_LAUNCH_TILES);
        int sub = 1;
        while (sub < 8 && (lhi - lo) * sub < sms) sub <<= 1;   // (splitting tiles further to fill whole waves measured slower:
                                                                // every unit re-stages its tile)
        const unsigned grid = static_cast<unsigned>(std::min<int64_t>((lhi - lo) * sub, sms));
        switch (sc.qs) {
#define ARB_MJ(QS_) case QS_: bm25_member_kernel<QS_><<<grid, MJ_THREADS, mj_smem, st>>>(ix, sc, lo, lhi, sub, mj_ch, dbg); \
                             bm25_member_exact_kernel<QS_><<<dim3(32, grid), 256, 0, st>>>(ix, sc); break;
          ARB_MJ(4) ARB_MJ(8) ARB_MJ(12) ARB_MJ(16) ARB_MJ(20) ARB_MJ(24) ARB_MJ(28) ARB_MJ(32)
#undef ARB_MJ
          default: set_error("bm25 batch: more than 32 tokens per query"); return ARB_ELIMIT;
        }
        ARB_LAUNCH_OK();
      }
      if (hi > prev || last) {
        bm25_compact_kernel<<<static_cast<unsigned>(nq), 256, CP_SMEM, st>>>(nq, k, last ? 1 : 0, sc);
        ARB_LAUNCH_OK();
      }
      prev = hi;
    }
  }
  if (planes) {
    ProfileScope ps(PROF_BM25_PLANES, st);
    ARB_CUDA(cudaMemsetAsync(sc.hist, 0, static_cast<size_t>(H3_MAX) * H3_BINS * 4, st));
    // ---- 2a. queries whose list A holds fewer than k docs although they have plane tokens have no bound for the screening
    // pass.  A histogram of the plane-only scores over a SAMPLE of the tiles (every 4th) gives one: the edge at which the
    // count from the top reaches k + the query's members is met by at least k non-member docs, whose complete score is
    // their plane-only score -- a valid lower bound of the final k-th score.  The query then takes the ordinary screening
    // pass with that threshold (about 4k emissions; an overflowing list E sends it to the per-query kernel).
    {
      if (nq > 2048) {
        bm25_h3_decide_kernel<<<1, 256, 0, st>>>(sc);
        ARB_LAUNCH_OK();
      }
      const int step = ix.n_tiles > 256 ? 4 : 1;
      const size_t plane_smem = static_cast<size_t>(ix.n_planes) * PLANE_TILE * 8;
      const unsigned h3_grid = static_cast<unsigned>(std::min<int64_t>((ix.n_tiles + step - 1) / step, sms));
      bm25_plane_h3_kernel<false><<<h3_grid, PL_WARPS * 32, plane_smem, st>>>(ix, sc, step);
      ARB_LAUNCH_OK();
      bm25_h3_threshold_kernel<<<H3_MAX / 4, 128, 0, st>>>(k, sc);
      ARB_LAUNCH_OK();
    }
    // ---- 2. screening pass: one CTA per SM; query groups sized by what is left of shared memory beside two tiles
    const size_t tiles_smem = 2 * static_cast<size_t>(ix.n_planes) * PLANE_TILE * 4 + 2 * static_cast<size_t>((ix.n_planes + 1) & ~1) * 8;
    const int64_t max_qpg = static_cast<int64_t>((SMEM_MAX - tiles_smem) / 20);     // {thr, slots} record + query index per query
    int64_t groups = std::max<int64_t>(1, (nq + max_qpg - 1) / max_qpg);
    if (ix.n_tiles * groups < sms)                                                  // few tiles: more groups fill the GPU
      groups = std::max<int64_t>(groups, std::min<int64_t>((sms + ix.n_tiles - 1) / ix.n_tiles, (nq + 4 * SC_WARPS - 1) / (4 * SC_WARPS)));
    const int64_t qpg = (nq + groups - 1) / groups;
    groups = (nq + qpg - 1) / qpg;
    const int64_t sets = std::max<int64_t>(1, std::min<int64_t>(ix.n_tiles, sms / groups));
    dim3 grid(static_cast<unsigned>(sets), static_cast<unsigned>(groups));
    static const int grab = std::getenv("ARB_BM25_GRAB") ? std::max(1, std::min(32, atoi(std::getenv("ARB_BM25_GRAB")))) : SC_GRAB;
    bm25_screen_kernel<<<grid, SC_WARPS * 32, tiles_smem + static_cast<size_t>(qpg) * 20, st>>>(ix, nq, qpg, sc, grab);
    ARB_LAUNCH_OK();
    bm25_recheck_kernel<<<static_cast<unsigned>(8 * sms), 256, 0, st>>>(ix, sc);
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
    std::vector<int> h(static_cast<size_t>(nq) + 8);
    ARB_CUDA(cudaMemcpyAsync(h.data(), sc.e_cnt, h.size() * 4, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaStreamSynchronize(st));
    g_stats[0] += nq;
    g_stats[1] += h[nq + 1];
    g_stats[4] += h[nq + 3];
    g_stats[5] += h[nq + 0];
    g_stats[6] += h[nq + 4];
    for (int64_t i = 0; i < nq; ++i) { g_stats[2] += h[i]; g_stats[3] = std::max<int64_t>(g_stats[3], h[i]); }
  }
  return ARB_OK;
}

int bm25_batch_search(const Bm25Dev& ix, const int32_t* d_terms, int64_t nq, int max_qt, int k, int64_t id_base,
                      double* d_scores, int64_t* d_ids, cudaStream_t st) {
  int rc = batch_set_attrs();
  if (rc) return rc;
  // candidate lists are C_CAP entries per query: cut very large batches so that the scratch stays under ~0.4 GB
  int64_t chunk = 8192;
  if (const char* v = std::getenv("ARB_BM25_CHUNK")) chunk = std::max<int64_t>(1, atoll(v));
  for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
    const int64_t n = std::min(chunk, nq - q0);
    rc = batch_chunk(ix, d_terms + q0 * max_qt, n, max_qt, k, id_base, d_scores + q0 * k, d_ids + q0 * k, st);
    if (rc) return rc;
  }
  return ARB_OK;
}

}  // namespace arb

extern "C" int arb_bm25_batch_stats(int64_t out[8], int reset) {
  for (int i = 0; i < 8; ++i) { if (out) out[i] = arb::g_stats[i]; if (reset) arb::g_stats[i] = 0; }
  return ARB_OK;
}
