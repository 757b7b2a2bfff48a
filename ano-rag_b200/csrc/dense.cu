// Dense exact inner-product top-k for B200 (sm_100a).
//
// Replaces faiss.IndexFlatIP.search as called at reference
// vector_store/vector_index.py:223 (and the row pre-processing :265-282).
//
//   scores[q, n] = sum over "terms" of  Qplane_a[q, :] . Xplane_b[n, :]
//
// runs as a persistent, warp-specialised tcgen05 GEMM (bf16 operands, fp32
// accumulators in TMEM, operands staged by TMA with the 128B swizzle) whose
// epilogue never writes the score block: each epilogue thread owns one query row
// (one TMEM lane), compares every score with that query's running threshold and
// appends the rare survivors to a small per-(work item, query) candidate list;
// when a list fills, the warp selects its k best in registers and publishes the
// k-th score as the new threshold (shared between all CTAs working on that query
// through one atomicMax).  A second kernel merges the lists to the final top-k.
//
// fp32-exact mode: rows are stored as three bf16 planes (hi/mid/lo, exact split of
// the fp32 value); the six largest cross terms are accumulated in fp32 on the
// tensor cores, smallest first.
#include "common.cuh"

#include <cub/cub.cuh>
#include "ptx.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>
#include <math.h>
#include <string.h>

namespace arb {

using namespace ptx;

constexpr int BLOCK_M = 128;              // queries per tile (= TMEM lanes)
constexpr int BLOCK_N = 256;              // corpus rows per tile (= UMMA N)
constexpr int BLOCK_K = 64;               // bf16 per k-block: one 128 B swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;   // 512: the whole TMEM of the SM
constexpr int EPI_HALVES = 2;                     // epilogue warps per TMEM lane quadrant: each takes half of a tile's columns
constexpr int DENSE_THREADS = 64 + 128 * EPI_HALVES;   // warp0 TMA, warp1 MMA, warps 2..9 epilogue
constexpr int DENSE_SMEM = STAGES * STAGE_BYTES + 256 + 1024;
constexpr int PSTAGES = 6;                        // CTA-pair kernel: A 128 x 64 + half of B (128 x 64) per stage
constexpr int PSTAGE_BYTES = A_BYTES + B_BYTES / 2;
constexpr int PAIR_SMEM = PSTAGES * PSTAGE_BYTES + 256 + 1024;
constexpr int MAX_TERMS = 6;

constexpr uint32_t KEY_NEG_INF = 0x007FFFFFu;     // f32_key(-inf)

struct IvfItem { uint32_t row_lo, row_hi; };   // rows of the inverted list this A-tile (item i = gathered rows [128i, 128i+128)) scans

struct DenseParams {
  int nq;
  uint32_t n_rows;
  int k;
  int kb_per_term;          // d_pad / 64
  int n_terms;
  int a_col[MAX_TERMS];     // element offset of the plane inside a query row
  int b_col[MAX_TERMS];     // element offset of the plane inside a corpus row
  int n_qtiles, n_splits, tiles_per_split, n_ntiles, n_items;
  uint2* lists;             // [n_items, 128, C] {row index, ordered score key}
  int* counts;              // [n_items, 128]
  uint32_t* gthr;           // [n_qtiles*128] shared per-query threshold key
  // Cross-list bound (flat scan, small batches): a query's rows are split over n_lists = n_splits * EPI_HALVES
  // candidate lists.  Every list tracks the rank2-th best key it has met (rank2 = ceil(k / n_lists) <= 8) and
  // publishes it once per tile; as soon as all lists have one, at least n_lists * rank2 >= k scores reach the
  // smallest published key, so the query's threshold can be raised to it long before any single list has seen k
  // good rows.  Without it a list of a 512-query batch (74 lists per query) never gets a selective threshold.
  uint32_t* sthr;           // [n_lists, n_qtiles*128] published keys, 0 = none yet; NULL = off
  int n_lists, rank2;
  // IVF list scan (template flag IVF): work items come from a device-built table
  const IvfItem* items;     // [n_items] A-tile of gathered queries x one inverted list
  const int* n_items_ptr;   // number of items (known only on the device)
  const int* row_query;     // [pair rows] query of each gathered row, -1 = padding
  const int* pair_pos;      // [nq, nprobe] gathered-row position of each (query, probe) pair
  int nprobe;
  const int64_t* row_ids;   // [n_rows] original id of each stored (list-grouped) row
};

// ----------------------------------------------------------------------------
// Warp-cooperative selection: keep the k largest of list[0..cnt) (cnt <= 32*EPL)
// in list[0..k), return the k-th largest key.  Exact ties at the k-th key are
// resolved towards the smaller row index.  All 32 lanes participate.
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint2 ld_volatile_u2(const uint2* p) {
  uint2 v;
  const volatile uint32_t* q = reinterpret_cast<const volatile uint32_t*>(p);
  v.x = q[0];
  v.y = q[1];
  return v;
}

template <int EPL, bool GLOBAL>
__device__ __forceinline__ uint32_t warp_select(uint2* list, int cnt, int k, int lane) {
  uint32_t key[EPL], idx[EPL];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    int p = lane + 32 * i;
    uint2 e = make_uint2(0u, 0u);
    if (p < cnt) { if (GLOBAL) e = __ldcg(list + p); else e = ld_volatile_u2(list + p); }
    idx[i] = e.x;
    key[i] = e.y;
  }
  uint32_t T = 0;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    uint32_t cand = T | (1u << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < EPL; ++i) c += (key[i] >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) T = cand;
  }
  const uint32_t lt = (1u << lane) - 1u;
  int base = 0;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    bool gt = key[i] > T;
    uint32_t m = __ballot_sync(0xffffffffu, gt);
    if (gt) list[base + __popc(m & lt)] = make_uint2(idx[i], key[i]);
    base += __popc(m);
  }
  const int need = k - base;
  // entries equal to the k-th key: keep the `need` smallest row indices, so the survivors are the top-k by
  // (key descending, index ascending) whatever order the list was filled in
  int ties = 0;
#pragma unroll
  for (int i = 0; i < EPL; ++i) ties += ((key[i] == T) && (T != 0u)) ? 1 : 0;
  ties = __reduce_add_sync(0xffffffffu, ties);
  uint32_t imax = 0xFFFFFFFFu;                       // largest index that is kept
  if (ties > need) {
    imax = 0;                                        // -> the need-th smallest index among the ties
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = imax | (1u << bit);
      int c = 0;
#pragma unroll
      for (int i = 0; i < EPL; ++i) c += (key[i] == T && idx[i] < cand) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c < need) imax = cand;
    }
  }
  int tbase = 0;
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    bool eq = (key[i] == T) && (T != 0u) && idx[i] <= imax;
    uint32_t m = __ballot_sync(0xffffffffu, eq);
    int r = tbase + __popc(m & lt);
    if (eq && r < need) list[base + r] = make_uint2(idx[i], key[i]);
    tbase += __popc(m);
  }
  __syncwarp();
  return T;
}

// ----------------------------------------------------------------------------
// The fused GEMM + threshold-select kernel.
// ----------------------------------------------------------------------------
template <int EPL, bool IVF, int CROSS = 0>      // CROSS: 0 = off, else how many best keys a list tracks (4 or 8)
__global__ void __launch_bounds__(DENSE_THREADS, 1)
dense_topk_kernel(const __grid_constant__ CUtensorMap tmap_a,
                  const __grid_constant__ CUtensorMap tmap_b, const DenseParams p) {
  constexpr int C = 32 * EPL;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4 * EPI_HALVES); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int kb_total = p.n_terms * p.kb_per_term;
  const int n_items = IVF ? __ldg(p.n_items_ptr) : p.n_items;
  // work item -> (A tile row, first B row, n-tile range, end of the valid B rows)
  auto decode = [&](int item, int& a_row0, uint32_t& b_row0, int& t0, int& t1, uint32_t& b_end) {
    if (IVF) {
      const IvfItem it = p.items[item];
      a_row0 = item * BLOCK_M; b_row0 = it.row_lo; b_end = it.row_hi;
      t0 = 0; t1 = static_cast<int>((it.row_hi - it.row_lo + BLOCK_N - 1) / BLOCK_N);
    } else {
      const int qt = item % p.n_qtiles, sp = item / p.n_qtiles;
      a_row0 = qt * BLOCK_M; b_row0 = 0; b_end = p.n_rows;
      t0 = sp * p.tiles_per_split; t1 = min(t0 + p.tiles_per_split, p.n_ntiles);
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int a_row0, t0, t1; uint32_t b_row0, b_end;
        decode(item, a_row0, b_row0, t0, t1, b_end);
        for (int t = t0; t < t1; ++t) {
          for (int term = 0; term < p.n_terms; ++term) {
            for (int kb = 0; kb < p.kb_per_term; ++kb) {
              mbar_wait(&empty[stage], phase ^ 1u);
              uint8_t* sa = smem + stage * STAGE_BYTES;
              mbar_expect_tx(&full[stage], STAGE_BYTES);
              tma_load_2d(&tmap_a, &full[stage], sa, p.a_col[term] + kb * BLOCK_K, a_row0, EVICT_LAST);
              tma_load_2d(&tmap_b, &full[stage], sa + A_BYTES, p.b_col[term] + kb * BLOCK_K,
                          static_cast<int32_t>(b_row0) + t * BLOCK_N, EVICT_NORMAL);
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int a_row0, t0, t1; uint32_t b_row0, b_end;
        decode(item, a_row0, b_row0, t0, t1, b_end);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&tempty[acc], acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
          for (int kb = 0; kb < kb_total; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t a_desc = umma_desc_k_sw128(sa);
            const uint64_t b_desc = umma_desc_k_sw128(sa + A_BYTES);
#pragma unroll
            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
              // +32 B per UMMA_K inside the 128 B swizzle row -> +2 in 16 B units
              umma_bf16(d_tmem, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (kb | kk) != 0 ? 1u : 0u);
            }
            umma_commit(&empty[stage]);          // frees the smem slot when these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          umma_commit(&tfull[acc]);              // accumulator complete -> epilogue
          if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (4 * EPI_HALVES warps)
    const int quad = warp & 3;                   // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;            // which half of every tile's columns this warp scans
    const int row = quad * 32 + lane;            // query row inside the tile
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int a_row0, t0, t1; uint32_t b_row0, b_end;
      decode(item, a_row0, b_row0, t0, t1, b_end);
      const int q = IVF ? __ldg(p.row_query + a_row0 + row) : a_row0 + row;
      const bool valid = q >= 0 && q < p.nq;
      uint2* list = p.lists + ((static_cast<size_t>(item) * BLOCK_M + row) * EPI_HALVES + half) * C;
      int cnt = 0;
      uint32_t thr_key = KEY_NEG_INF;
      float thr = valid ? -INFINITY : INFINITY;
      constexpr bool cross = CROSS > 0 && !IVF;
      constexpr int NB = CROSS > 0 ? CROSS : 1;
      const size_t q_pad = static_cast<size_t>(p.n_qtiles) * BLOCK_M;
      const int li = IVF ? 0 : (item / p.n_qtiles) * EPI_HALVES + half;
      uint32_t best[NB], b_pub = 0;                         // the NB best keys this list has met, descending
#pragma unroll
      for (int i = 0; i < NB; ++i) best[i] = 0;
      for (int t = t0; t < t1; ++t) {
        if (cross && valid) {
          uint32_t br = best[NB - 1];
#pragma unroll
          for (int i = 0; i < NB - 1; ++i) if (p.rank2 == i + 1) br = best[i];
          if (br != b_pub) { b_pub = br; *reinterpret_cast<volatile uint32_t*>(p.sthr + li * q_pad + q) = br; }
          if ((t - t0) % p.n_lists == li % p.n_lists || t == t0 + 2) {   // the lists of a query take turns (warp-uniform)
            uint32_t mn = 0xFFFFFFFFu;
            for (int i = 0; i < p.n_lists; ++i) mn = min(mn, i == li ? br : __ldcg(p.sthr + i * q_pad + q));
            if (mn > KEY_NEG_INF + 1u) atomicMax(p.gthr + q, mn - 1u);   // "score > threshold" keeps scores equal to mn
          }
        }
        if (valid) {
          uint32_t g = __ldcg(p.gthr + q);
          if (g > thr_key) { thr_key = g; thr = key_f32(g); }
        }
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t n_base = b_row0 + static_cast<uint32_t>(t) * BLOCK_N;
        const bool tail = n_base + BLOCK_N > b_end;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
        for (int c = half * (BLOCK_N / 32 / EPI_HALVES); c < (half + 1) * (BLOCK_N / 32 / EPI_HALVES); ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          if (c == (half + 1) * (BLOCK_N / 32 / EPI_HALVES) - 1) {
            // all of this warp's reads of the accumulator are done: hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          const uint32_t idx0 = n_base + c * 32;
          if (!tail) {
            // Once the threshold has warmed up almost no score beats it: a max-reduction per group of 8 columns
            // decides whether that group's compare-and-append code runs at all (the epilogue warps, one per
            // scheduler, were the bottleneck of the kernel when every column paid for the append path).
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              float mx = __uint_as_float(v[8 * g8]);
#pragma unroll
              for (int j = 1; j < 8; ++j) mx = fmaxf(mx, __uint_as_float(v[8 * g8 + j]));
              if (mx > thr) {                                // taken by the whole warp if any of its 32 rows passes
#pragma unroll
                for (int j = 8 * g8; j < 8 * g8 + 8; ++j) {
                  float s = __uint_as_float(v[j]);
                  if (s > thr) {
                    const uint32_t key = f32_key(s);
                    list[cnt] = make_uint2(idx0 + j, key); ++cnt;
                    if (cross && key > best[NB - 1]) {       // sorted insert into the NB best
                      best[NB - 1] = key;
#pragma unroll
                      for (int i = NB - 1; i > 0; --i)
                        if (best[i] > best[i - 1]) { const uint32_t x = best[i - 1]; best[i - 1] = best[i]; best[i] = x; }
                    }
                  }
                }
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float s = __uint_as_float(v[j]);
              if (s > thr && idx0 + j < b_end) { list[cnt] = make_uint2(idx0 + j, f32_key(s)); ++cnt; }
            }
          }
          uint32_t need = __ballot_sync(0xffffffffu, cnt > C - 32);
          while (need) {
            const int r = __ffs(need) - 1;
            need &= need - 1;
            uint2* l = reinterpret_cast<uint2*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(list), r));
            const int c_r = __shfl_sync(0xffffffffu, cnt, r);
            const uint32_t T = warp_select<EPL, true>(l, c_r, p.k, lane);
            if (lane == r) {
              cnt = p.k;
              if (T > thr_key) { thr_key = T; thr = key_f32(T); }
              atomicMax(p.gthr + (valid ? q : 0), T);
            }
          }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      }
      p.counts[(static_cast<size_t>(item) * BLOCK_M + row) * EPI_HALVES + half] = valid ? cnt : 0;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// CTA-pair variant (cta_group::2): a cluster of two CTAs owns one 256-query x 256-row tile.  Each CTA stages its own
// 128 query rows and HALF of the corpus tile (128 rows) per k-block, so a stage is 32 KB instead of 48 KB and six
// stages fit: 1.5x more k-blocks in flight for the same shared memory, and a third less L2 -> SM traffic per flop.
// The leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256); the accumulator rows of each CTA's
// queries land in that CTA's own TMEM, so the epilogue is the one of the single-CTA kernel.
template <int EPL>
__global__ void __launch_bounds__(DENSE_THREADS, 1)
dense_topk_pair_kernel(const __grid_constant__ CUtensorMap tmap_a,
                  const __grid_constant__ CUtensorMap tmap_b, const DenseParams p) {
  constexpr bool IVF = false;
  constexpr int C = 32 * EPL;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PSTAGES * PSTAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + PSTAGES;
  uint64_t* tfull = bars + 2 * PSTAGES;
  uint64_t* tempty = tfull + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < PSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    // the leader's "accumulator drained" barrier collects the epilogue warps of BOTH CTAs
    for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8 * EPI_HALVES); }
    fence_barrier_init();
  }
  __syncthreads();
  cluster_sync_all();                              // barriers of both CTAs exist before anything remote happens
  if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int kb_total = p.n_terms * p.kb_per_term;
  const int n_items = IVF ? __ldg(p.n_items_ptr) : p.n_items;
  // work item -> (A tile row, first B row, n-tile range, end of the valid B rows)
  auto decode = [&](int item, int& a_row0, uint32_t& b_row0, int& t0, int& t1, uint32_t& b_end) {
    if (IVF) {
      const IvfItem it = p.items[item];
      a_row0 = item * BLOCK_M; b_row0 = it.row_lo; b_end = it.row_hi;
      t0 = 0; t1 = static_cast<int>((it.row_hi - it.row_lo + BLOCK_N - 1) / BLOCK_N);
    } else {
      const int n_qt2 = p.n_qtiles >> 1;             // p.n_qtiles is even in pair mode
      const int qt = 2 * (item % n_qt2) + static_cast<int>(cta_rank), sp = item / n_qt2;
      a_row0 = qt * BLOCK_M; b_row0 = 0; b_end = p.n_rows;
      t0 = sp * p.tiles_per_split; t1 = min(t0 + p.tiles_per_split, p.n_ntiles);
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = pair; item < n_items; item += n_pairs) {
        int a_row0, t0, t1; uint32_t b_row0, b_end;
        decode(item, a_row0, b_row0, t0, t1, b_end);
        for (int t = t0; t < t1; ++t) {
          for (int term = 0; term < p.n_terms; ++term) {
            for (int kb = 0; kb < p.kb_per_term; ++kb) {
              mbar_wait(&empty[stage], phase ^ 1u);
              uint8_t* sa = smem + stage * PSTAGE_BYTES;
              // both CTAs' bytes are accounted on the leader's barrier (its arrival carries the expected total)
              if (leader) mbar_expect_tx(&full[stage], 2 * PSTAGE_BYTES);
              tma_load_2d_2sm(&tmap_a, &full[stage], sa, p.a_col[term] + kb * BLOCK_K, a_row0, EVICT_LAST);
              tma_load_2d_2sm(&tmap_b, &full[stage], sa + A_BYTES, p.b_col[term] + kb * BLOCK_K,
                              static_cast<int32_t>(b_row0) + t * BLOCK_N + static_cast<int32_t>(cta_rank) * (BLOCK_N / 2),
                              EVICT_NORMAL);
              if (++stage == PSTAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BLOCK_M, BLOCK_N);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int item = pair; item < n_items; item += n_pairs) {
        int a_row0, t0, t1; uint32_t b_row0, b_end;
        decode(item, a_row0, b_row0, t0, t1, b_end);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&tempty[acc], acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
          for (int kb = 0; kb < kb_total; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * PSTAGE_BYTES);
            const uint64_t a_desc = umma_desc_k_sw128(sa);
            const uint64_t b_desc = umma_desc_k_sw128(sa + A_BYTES);
#pragma unroll
            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
              // +32 B per UMMA_K inside the 128 B swizzle row -> +2 in 16 B units
              umma_bf16_2sm(d_tmem, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (kb | kk) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(&empty[stage]);      // frees the slot in both CTAs when these MMAs retire
            if (++stage == PSTAGES) { stage = 0; phase ^= 1u; }
          }
          umma_commit_2sm(&tfull[acc]);          // accumulator complete -> the epilogues of both CTAs
          if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (4 * EPI_HALVES warps)
    const int quad = warp & 3;                   // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;            // which half of every tile's columns this warp scans
    const int row = quad * 32 + lane;            // query row inside the tile
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = pair; item < n_items; item += n_pairs) {
      int a_row0, t0, t1; uint32_t b_row0, b_end;
      decode(item, a_row0, b_row0, t0, t1, b_end);
      const int q = IVF ? __ldg(p.row_query + a_row0 + row) : a_row0 + row;
      const bool valid = q >= 0 && q < p.nq;
      // candidate-list slot of this CTA's query tile: same numbering as the single-CTA kernel (split-major)
      const int n_qt2 = p.n_qtiles >> 1;
      const size_t list_item = static_cast<size_t>(item / n_qt2) * p.n_qtiles + 2 * (item % n_qt2) + cta_rank;
      uint2* list = p.lists + ((list_item * BLOCK_M + row) * EPI_HALVES + half) * C;
      int cnt = 0;
      uint32_t thr_key = KEY_NEG_INF;
      float thr = valid ? -INFINITY : INFINITY;
      for (int t = t0; t < t1; ++t) {
        if (valid) {
          uint32_t g = __ldcg(p.gthr + q);
          if (g > thr_key) { thr_key = g; thr = key_f32(g); }
        }
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t n_base = b_row0 + static_cast<uint32_t>(t) * BLOCK_N;
        const bool tail = n_base + BLOCK_N > b_end;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
        for (int c = half * (BLOCK_N / 32 / EPI_HALVES); c < (half + 1) * (BLOCK_N / 32 / EPI_HALVES); ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          if (c == (half + 1) * (BLOCK_N / 32 / EPI_HALVES) - 1) {
            // all of this warp's reads of the accumulator are done: hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty[acc], 0));
          }
          const uint32_t idx0 = n_base + c * 32;
          if (!tail) {
            // Once the threshold has warmed up almost no score beats it: a max-reduction per group of 8 columns
            // decides whether that group's compare-and-append code runs at all (the epilogue warps, one per
            // scheduler, were the bottleneck of the kernel when every column paid for the append path).
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              float mx = __uint_as_float(v[8 * g8]);
#pragma unroll
              for (int j = 1; j < 8; ++j) mx = fmaxf(mx, __uint_as_float(v[8 * g8 + j]));
              if (mx > thr) {                                // taken by the whole warp if any of its 32 rows passes
#pragma unroll
                for (int j = 8 * g8; j < 8 * g8 + 8; ++j) {
                  float s = __uint_as_float(v[j]);
                  if (s > thr) { list[cnt] = make_uint2(idx0 + j, f32_key(s)); ++cnt; }
                }
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float s = __uint_as_float(v[j]);
              if (s > thr && idx0 + j < b_end) { list[cnt] = make_uint2(idx0 + j, f32_key(s)); ++cnt; }
            }
          }
          uint32_t need = __ballot_sync(0xffffffffu, cnt > C - 32);
          while (need) {
            const int r = __ffs(need) - 1;
            need &= need - 1;
            uint2* l = reinterpret_cast<uint2*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(list), r));
            const int c_r = __shfl_sync(0xffffffffu, cnt, r);
            const uint32_t T = warp_select<EPL, true>(l, c_r, p.k, lane);
            if (lane == r) {
              cnt = p.k;
              if (T > thr_key) { thr_key = T; thr = key_f32(T); }
              atomicMax(p.gthr + (valid ? q : 0), T);
            }
          }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      }
      p.counts[(list_item * BLOCK_M + row) * EPI_HALVES + half] = valid ? cnt : 0;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // the peer may still read this CTA's shared memory / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}


// warp_select for the merge kernel's shared-memory buffers as a real function: the merge kernel calls it from inside
// loops that are unrolled EPL times, and inlining it there (up to 32 copies per instantiation, nine instantiations)
// cost minutes of compile time for a path that runs once per overflowing list.
template <int EPL>
__device__ __noinline__ uint32_t warp_select_shared(uint2* list, int cnt, int k, int lane) {
  return warp_select<EPL, false>(list, cnt, k, lane);
}

// ----------------------------------------------------------------------------
// Merge the per-item candidate lists of each query into its final top-k.
// One CTA of MERGE_WARPS warps per query: the lists are dealt round-robin to the warps, every list is fetched
// whole (its count one list ahead) so a warp pays one memory round trip per list instead of one per 32 entries
// (a query has n_splits * EPI_HALVES lists -- 10 at 4096 queries, 74 at 512 -- and the serial version was
// latency-bound: ~1 ms of a 1.45 ms dense stage at 512 queries).  Warp 0 then folds the other warps' survivors.
// ----------------------------------------------------------------------------
// MERGE_WARPS = 4 for batches (few lists per query, thousands of CTAs), 16 for small batches: a single query against
// 1 M rows has 2 x 148 lists, and the merge was a third of its latency with four warps.
template <int EPL, bool IVF, int MERGE_WARPS = 4>
__global__ void __launch_bounds__(32 * MERGE_WARPS)
dense_merge_kernel(const DenseParams p, int64_t id_base, float* __restrict__ out_scores,
                   int64_t* __restrict__ out_ids) {
  constexpr int C = 32 * EPL;
  extern __shared__ uint8_t msmem[];
  __shared__ uint32_t s_thr;
  __shared__ int s_cnt[MERGE_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x;
  uint2* buf = reinterpret_cast<uint2*>(msmem) + static_cast<size_t>(warp) * C;
  const int qt = q / BLOCK_M, row = q % BLOCK_M;
  const uint32_t lt = (1u << lane) - 1u;
  if (threadIdx.x == 0) s_thr = p.gthr[q];
  __syncthreads();
  uint32_t thr_key = s_thr;
  int cnt = 0;
  const int n_lists = (IVF ? p.nprobe : p.n_splits) * EPI_HALVES;
  // list li -> its index in p.lists / p.counts, or -1 (IVF pair without a gathered row)
  auto list_of = [&](int li) -> long long {
    const int sp = li / EPI_HALVES, hb = li % EPI_HALVES;
    size_t slot;
    if (IVF) {
      const int pos = p.pair_pos[static_cast<size_t>(q) * p.nprobe + sp];
      if (pos < 0) return -1;
      slot = static_cast<size_t>(pos);
    } else {
      slot = (static_cast<size_t>(sp) * p.n_qtiles + qt) * BLOCK_M + row;
    }
    return static_cast<long long>(slot * EPI_HALVES + hb);
  };
  long long cur = warp < n_lists ? list_of(warp) : -1;
  int n_cur = cur >= 0 ? p.counts[cur] : 0;
  for (int li = warp; li < n_lists; li += MERGE_WARPS) {
    const long long nxt = li + MERGE_WARPS < n_lists ? list_of(li + MERGE_WARPS) : -1;
    const int n_nxt = nxt >= 0 ? p.counts[nxt] : 0;          // in flight while this list is merged
    const int n = n_cur;
    if (n > 0) {
      const uint2* L = p.lists + static_cast<size_t>(cur) * C;
      uint2 e[EPL];
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        e[i] = make_uint2(0u, 0u);
        if (32 * i + lane < n) e[i] = __ldcg(L + 32 * i + lane);
      }
      if (IVF) {
#pragma unroll
        for (int i = 0; i < EPL; ++i)      // rank and tie-break by the original row, not the list slot
          if (32 * i + lane < n) e[i].x = static_cast<uint32_t>(p.row_ids[e[i].x]);
      }
      thr_key = max(thr_key, *reinterpret_cast<volatile uint32_t*>(&s_thr));
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        if (32 * i < n) {                                     // warp-uniform
          const bool pass = (32 * i + lane < n) && e[i].y >= thr_key;
          const uint32_t m = __ballot_sync(0xffffffffu, pass);
          if (pass) buf[cnt + __popc(m & lt)] = e[i];
          cnt += __popc(m);
          if (cnt > C - 32) {
            const uint32_t T = warp_select_shared<EPL>(buf, cnt, p.k, lane);
            cnt = p.k;
            thr_key = max(thr_key, T);
            if (lane == 0) atomicMax(&s_thr, T);
          }
        }
      }
    }
    cur = nxt; n_cur = n_nxt;
  }
  __syncwarp();
  if (cnt > p.k) { warp_select_shared<EPL>(buf, cnt, p.k, lane); cnt = p.k; }
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  if (warp != 0) return;
  // fold the survivors of warps 1.. into warp 0's buffer (each holds at most k <= C - 32 entries)
  thr_key = max(thr_key, *reinterpret_cast<volatile uint32_t*>(&s_thr));
  for (int w = 1; w < MERGE_WARPS; ++w) {
    const uint2* B = reinterpret_cast<const uint2*>(msmem) + static_cast<size_t>(w) * C;
    const int n = s_cnt[w];
    for (int b = 0; b < n; b += 32) {
      uint2 e = make_uint2(0u, 0u);
      if (b + lane < n) e = B[b + lane];
      const bool pass = (b + lane < n) && e.y >= thr_key;
      const uint32_t m = __ballot_sync(0xffffffffu, pass);
      if (pass) buf[cnt + __popc(m & lt)] = e;
      cnt += __popc(m);
      if (cnt > C - 32) {
        const uint32_t T = warp_select_shared<EPL>(buf, cnt, p.k, lane);
        cnt = p.k;
        thr_key = max(thr_key, T);
      }
    }
  }
  __syncwarp();
  if (cnt > p.k) { warp_select_shared<EPL>(buf, cnt, p.k, lane); cnt = p.k; }
  __syncwarp();
  // rank by (score desc, id asc): cnt <= k <= 256, O(cnt^2 / 32) per lane
  for (int j = lane; j < cnt; j += 32) {
    const uint2 e = buf[j];
    int rank = 0;
    for (int i = 0; i < cnt; ++i) {
      const uint2 o = ld_volatile_u2(buf + i);
      rank += (o.y > e.y || (o.y == e.y && o.x < e.x)) ? 1 : 0;
    }
    out_scores[static_cast<size_t>(q) * p.k + rank] = key_f32(e.y);
    out_ids[static_cast<size_t>(q) * p.k + rank] = id_base + static_cast<int64_t>(e.x);
  }
  for (int j = cnt + lane; j < p.k; j += 32) {
    out_scores[static_cast<size_t>(q) * p.k + j] = -3.4028234663852886e38f;   // faiss "no result"
    out_ids[static_cast<size_t>(q) * p.k + j] = -1;
  }
}

__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, size_t n) {
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}

// ----------------------------------------------------------------------------
// Row pre-processing (vector_index.py:265-282): fp32 cast, optional L2
// normalisation (norm == 0 -> 1), then bf16 planes.  One warp per row.
//   planes == 1: round-to-nearest-even bf16
//   planes == 3: exact truncation split hi/mid/lo (hi+mid+lo == fp32 value)
// Output row layout: [plane0 d_pad | plane1 d_pad | plane2 d_pad], zero padded.
// ----------------------------------------------------------------------------
template <typename Tin>
__device__ __forceinline__ float load_f32(const Tin* p);
template <> __device__ __forceinline__ float load_f32<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float load_f32<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename Tin>
__global__ void __launch_bounds__(256)
prep_rows_kernel(const Tin* __restrict__ in, int64_t n, int d, int d_pad, int planes, int normalize,
                 __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const Tin* src = in + r * d;
  float inv_norm_den = 1.0f;
  if (normalize) {
    float ss = 0.f;
    for (int i = lane; i < d; i += 32) { float x = load_f32(src + i); ss = fmaf(x, x, ss); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    float nrm = sqrtf(ss);
    inv_norm_den = (nrm == 0.f) ? 1.0f : nrm;
  }
  __nv_bfloat16* dst = out + r * static_cast<int64_t>(planes) * d_pad;
  for (int i = lane; i < d_pad; i += 32) {
    float x = 0.f;
    if (i < d) {
      x = load_f32(src + i);
      if (normalize) x = __fdiv_rn(x, inv_norm_den);
    }
    if (planes == 1) {
      dst[i] = __float2bfloat16_rn(x);
    } else {
      float rem = x;
      for (int pl = 0; pl < planes; ++pl) {
        float hi = __uint_as_float(__float_as_uint(rem) & 0xFFFF0000u);
        dst[pl * d_pad + i] = __ushort_as_bfloat16(static_cast<unsigned short>(__float_as_uint(hi) >> 16));
        rem = __fsub_rn(rem, hi);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
reconstruct_rows_kernel(const __nv_bfloat16* __restrict__ rows, int64_t n, int d, int d_pad, int planes,
                        float* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n * d) return;
  const int64_t r = i / d;
  const int c = static_cast<int>(i % d);
  const __nv_bfloat16* src = rows + r * static_cast<int64_t>(planes) * d_pad;
  float acc = 0.f;
  for (int pl = planes - 1; pl >= 0; --pl) acc = __fadd_rn(acc, __bfloat162float(src[pl * d_pad + c]));
  out[i] = acc;
}

// ----------------------------------------------------------------------------
// Host side
// ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// bf16 matrix [rows, cols] row-major (pitch = cols), box {64, box_rows}, 128B swizzle.
static int make_tmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  ARB_CHECK(fn != nullptr, ARB_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
    return ARB_ECUDA;
  }
  return ARB_OK;
}

}  // namespace arb

using namespace arb;

// a mutex that a struct copy does not carry along (k-means borrows an index's rows through a by-value view)
struct HandleMutex {
  std::mutex m;
  HandleMutex() = default;
  HandleMutex(const HandleMutex&) {}
  HandleMutex& operator=(const HandleMutex&) { return *this; }
};

struct arb_index {
  int dim = 0, d_pad = 0, storage = 0, planes = 1, normalize = 1;
  int normalize_rows = 1;        // rows of the next add calls (arb_index_set_row_normalization); queries follow `normalize`
  int device = 0;
  int64_t n = 0, cap = 0;
  __nv_bfloat16* rows = nullptr;       // [cap, planes*d_pad]
  size_t row_elems() const { return static_cast<size_t>(planes) * d_pad; }
  // IVF (IndexIVFFlat, inner-product coarse quantiser; vector_index.py:85-91): nlist > 0 once centroids exist
  int nlist = 0;
  arb_index* coarse = nullptr;         // the centroids as a small flat index (same storage mode, no normalisation)
  float* centroids = nullptr;          // [nlist, dim] fp32
  int32_t* assign = nullptr;           // [cap] list of each row (add order)
  bool grouped_dirty = true;
  int64_t g_n = 0, g_cap = 0;
  __nv_bfloat16* g_rows = nullptr;     // rows regrouped list-major
  int64_t* g_ids = nullptr;            // [g_n] add-order position of each regrouped row
  uint32_t* list_off = nullptr;        // [nlist + 1]
  // Searches are re-entrant on a read-only index, but the first IVF search after an add regroups the rows (frees and
  // reallocates g_rows / g_ids): that step, and everything that writes the index, is serialised per handle.
  HandleMutex mu;
};

static int launch_prep(const void* src_dev, int dtype, int64_t n, int d, int d_pad, int planes, int normalize,
                       __nv_bfloat16* dst, cudaStream_t st) {
  if (n == 0) return ARB_OK;
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((n + wpb - 1) / wpb);
  if (dtype == ARB_F32)
    prep_rows_kernel<float><<<grid, wpb * 32, 0, st>>>(static_cast<const float*>(src_dev), n, d, d_pad, planes, normalize, dst);
  else
    prep_rows_kernel<__nv_bfloat16><<<grid, wpb * 32, 0, st>>>(static_cast<const __nv_bfloat16*>(src_dev), n, d, d_pad, planes, normalize, dst);
  ARB_LAUNCH_OK();
  return ARB_OK;
}

static int ivf_assign_rows(arb_index* ix, int64_t row0, int64_t n, cudaStream_t st);

extern "C" int arb_index_set_row_normalization(arb_index* ix, int on) {
  ARB_CHECK(ix, ARB_EINVAL, "arb_index_set_row_normalization: NULL index");
  ix->normalize_rows = on ? 1 : 0;
  return ARB_OK;
}

extern "C" int arb_index_create(arb_index** out, int dim, int storage, int normalize) {
  ARB_CHECK(out != nullptr, ARB_EINVAL, "arb_index_create: out is NULL");
  ARB_CHECK(dim > 0 && dim <= 8192, ARB_EINVAL, "arb_index_create: dim must be in [1, 8192]");
  ARB_CHECK(storage == ARB_STORE_BF16 || storage == ARB_STORE_FP32_SPLIT, ARB_EINVAL, "arb_index_create: unknown storage mode");
  int dev = 0;
  ARB_CUDA(cudaGetDevice(&dev));
  arb_index* ix = new arb_index();
  ix->dim = dim;
  ix->d_pad = (dim + BLOCK_K - 1) / BLOCK_K * BLOCK_K;
  ix->storage = storage;
  ix->planes = storage == ARB_STORE_BF16 ? 1 : 3;
  ix->normalize = normalize ? 1 : 0;
  ix->normalize_rows = ix->normalize;
  ix->device = dev;
  *out = ix;
  return ARB_OK;
}

extern "C" int arb_index_destroy(arb_index* ix) {
  if (!ix) return ARB_OK;
  if (ix->rows) cudaFree(ix->rows);
  if (ix->coarse) arb_index_destroy(ix->coarse);
  if (ix->centroids) cudaFree(ix->centroids);
  if (ix->assign) cudaFree(ix->assign);
  if (ix->g_rows) cudaFree(ix->g_rows);
  if (ix->g_ids) cudaFree(ix->g_ids);
  if (ix->list_off) cudaFree(ix->list_off);
  delete ix;
  return ARB_OK;
}

extern "C" int arb_index_reset(arb_index* ix) {
  ARB_CHECK(ix != nullptr, ARB_EINVAL, "arb_index_reset: NULL index");
  ix->n = 0;
  ix->g_n = 0;
  ix->grouped_dirty = true;
  return ARB_OK;
}

extern "C" int64_t arb_index_ntotal(const arb_index* ix) { return ix ? ix->n : 0; }
extern "C" int arb_index_dim(const arb_index* ix) { return ix ? ix->dim : 0; }

extern "C" int arb_index_add(arb_index* ix, const void* rows, int dtype, int64_t n, int rows_is_device, void* stream) {
  ARB_CHECK(ix != nullptr, ARB_EINVAL, "arb_index_add: NULL index");
  ARB_CHECK(n >= 0 && (rows != nullptr || n == 0), ARB_EINVAL, "arb_index_add: bad rows/n");
  ARB_CHECK(dtype == ARB_F32 || dtype == ARB_BF16, ARB_EINVAL, "arb_index_add: dtype must be ARB_F32 or ARB_BF16");
  ARB_CHECK(ix->n + n <= 0xFFFFFFF0ll, ARB_ELIMIT, "arb_index_add: more than 2^32 rows per shard");
  if (n == 0) return ARB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ix->n + n > ix->cap) {
    int64_t ncap = ix->cap + ix->cap / 2;
    if (ncap < ix->n + n) ncap = ix->n + n;
    __nv_bfloat16* nr = nullptr;
    cudaError_t e = cudaMalloc(&nr, static_cast<size_t>(ncap) * ix->row_elems() * 2);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("arb_index_add: out of device memory"); return ARB_ENOMEM; }
    if (ix->n > 0) {
      ARB_CUDA(cudaMemcpyAsync(nr, ix->rows, static_cast<size_t>(ix->n) * ix->row_elems() * 2, cudaMemcpyDeviceToDevice, st));
      ARB_CUDA(cudaStreamSynchronize(st));
    }
    if (ix->rows) cudaFree(ix->rows);
    ix->rows = nr;
    int32_t* na = nullptr;
    e = cudaMalloc(&na, static_cast<size_t>(ncap) * 4);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("arb_index_add: out of device memory"); return ARB_ENOMEM; }
    if (ix->n > 0 && ix->assign) {
      ARB_CUDA(cudaMemcpyAsync(na, ix->assign, static_cast<size_t>(ix->n) * 4, cudaMemcpyDeviceToDevice, st));
      ARB_CUDA(cudaStreamSynchronize(st));
    }
    if (ix->assign) cudaFree(ix->assign);
    ix->assign = na;
    ix->cap = ncap;
  }
  const size_t esz = dtype == ARB_F32 ? 4 : 2;
  __nv_bfloat16* dst = ix->rows + static_cast<size_t>(ix->n) * ix->row_elems();
  if (rows_is_device) {
    int rc = launch_prep(rows, dtype, n, ix->dim, ix->d_pad, ix->planes, ix->normalize_rows, dst, st);
    if (rc) return rc;
  } else {
    // stage host rows through a bounded device buffer
    const int64_t chunk = std::max<int64_t>(1, (int64_t(256) << 20) / (ix->dim * esz));
    AsyncBuf stagebuf;
    ARB_CUDA(stagebuf.alloc(static_cast<size_t>(std::min(chunk, n)) * ix->dim * esz, st));
    for (int64_t off = 0; off < n; off += chunk) {
      const int64_t m = std::min(chunk, n - off);
      ARB_CUDA(cudaMemcpyAsync(stagebuf.p, static_cast<const char*>(rows) + off * ix->dim * esz,
                               static_cast<size_t>(m) * ix->dim * esz, cudaMemcpyHostToDevice, st));
      int rc = launch_prep(stagebuf.p, dtype, m, ix->dim, ix->d_pad, ix->planes, ix->normalize_rows,
                           dst + static_cast<size_t>(off) * ix->row_elems(), st);
      if (rc) return rc;
    }
    ARB_CUDA(cudaStreamSynchronize(st));
  }
  if (ix->nlist > 0) {                                   // IVF: quantise the new rows (add_with_ids, :187-193)
    int rc = ivf_assign_rows(ix, ix->n, n, st);
    if (rc) return rc;
  }
  ix->n += n;
  ix->grouped_dirty = true;
  return ARB_OK;
}

extern "C" int arb_index_reconstruct(const arb_index* ix, int64_t row0, int64_t n, float* out, int out_is_device, void* stream) {
  ARB_CHECK(ix != nullptr && out != nullptr, ARB_EINVAL, "arb_index_reconstruct: NULL argument");
  ARB_CHECK(row0 >= 0 && n >= 0 && row0 + n <= ix->n, ARB_EINVAL, "arb_index_reconstruct: row range outside the index");
  if (n == 0) return ARB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AsyncBuf tmp;
  float* dev_out = out;
  if (!out_is_device) { ARB_CUDA(tmp.alloc(static_cast<size_t>(n) * ix->dim * 4, st)); dev_out = tmp.as<float>(); }
  const int64_t total = n * ix->dim;
  reconstruct_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      ix->rows + static_cast<size_t>(row0) * ix->row_elems(), n, ix->dim, ix->d_pad, ix->planes, dev_out);
  ARB_LAUNCH_OK();
  if (!out_is_device) {
    ARB_CUDA(cudaMemcpyAsync(out, dev_out, static_cast<size_t>(total) * 4, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaStreamSynchronize(st));
  }
  return ARB_OK;
}

// choose how many corpus splits each query tile is cut into so that the persistent
// grid finishes in the fewest n-tile steps (ties: fewer splits = fewer lists)
static void plan_splits(int n_qtiles, int n_ntiles, int grid, int* n_splits, int* tiles_per_split, int* n_rounds = nullptr) {
  long best_cost = -1; int best_s = 1; long best_rounds = 1;
  long one_cost = -1; int one_s = 1;                    // best plan whose items all run at once (one round)
  const int max_s = std::max(1, std::min(n_ntiles, 64 * std::max(1, grid / std::max(1, n_qtiles)) + 8));
  for (int s = 1; s <= max_s; ++s) {
    const int tps = (n_ntiles + s - 1) / s;
    const int s_eff = (n_ntiles + tps - 1) / tps;
    const long items = static_cast<long>(s_eff) * n_qtiles;
    const long rounds = (items + grid - 1) / grid;
    const long cost = rounds * tps * 64 + s_eff;       // makespan first, then fewer lists
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_s = s_eff; best_rounds = rounds; }
    if (rounds == 1 && (one_cost < 0 || cost < one_cost)) { one_cost = cost; one_s = s_eff; }
  }
  // All lists of a query being live at the same time lets them share one tight threshold (DenseParams::sthr), which is
  // worth more than a few percent of makespan: take the one-round plan unless it is clearly longer.
  if (n_rounds && one_cost >= 0 && one_cost <= best_cost + best_cost / 8) { best_s = one_s; best_rounds = 1; }
  *n_splits = best_s;
  *tiles_per_split = (n_ntiles + best_s - 1) / best_s;
  *n_splits = (n_ntiles + *tiles_per_split - 1) / *tiles_per_split;
  if (n_rounds) *n_rounds = static_cast<int>(best_rounds);
}

// (query plane, corpus plane) pairs, smallest magnitude first
static void setup_terms(DenseParams& p, int q_planes, int planes, int d_pad) {
  int terms[MAX_TERMS][2]; int nt = 0;
  if (q_planes == 1 && planes == 1) { terms[nt][0] = 0; terms[nt++][1] = 0; }
  else if (q_planes == 3 && planes == 1) { for (int a = 2; a >= 0; --a) { terms[nt][0] = a; terms[nt++][1] = 0; } }
  else if (q_planes == 1 && planes == 3) { for (int b = 2; b >= 0; --b) { terms[nt][0] = 0; terms[nt++][1] = b; } }
  else {
    const int order[6][2] = {{2, 0}, {0, 2}, {1, 1}, {1, 0}, {0, 1}, {0, 0}};
    for (int i = 0; i < 6; ++i) { terms[nt][0] = order[i][0]; terms[nt++][1] = order[i][1]; }
  }
  p.n_terms = nt;
  for (int i = 0; i < nt; ++i) { p.a_col[i] = terms[i][0] * d_pad; p.b_col[i] = terms[i][1] * d_pad; }
}

// gthr[i] = thr_init[i] for real queries (pre-pass bound), "-inf" for padding rows
__global__ void init_gthr_kernel(uint32_t* __restrict__ gthr, const uint32_t* __restrict__ thr_init, int nq, size_t n) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) gthr[i] = (thr_init && i < static_cast<size_t>(nq)) ? thr_init[i] : KEY_NEG_INF;
}
// bound for the main pass from the pre-pass result: one key below its k-th score, so that the strict
// "score > threshold" test of the epilogue still admits that row when the main pass meets it again
__global__ void prepass_bound_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ids, int nq, int k,
                                     uint32_t* __restrict__ thr) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  thr[q] = ids[static_cast<size_t>(q) * k + k - 1] >= 0 ? f32_key(scores[static_cast<size_t>(q) * k + k - 1]) - 1u : KEY_NEG_INF;
}

// one scan of rows [0, n_rows) of the index; thr_init (device, [nq] keys or NULL) pre-loads the shared thresholds
template <int EPL>
static int run_dense_chunk(const arb_index* ix, int64_t n_rows, const uint32_t* thr_init, const __nv_bfloat16* qplanes,
                           int q_planes, int64_t nq, int k, int64_t id_base, float* d_scores, int64_t* d_ids,
                           cudaStream_t st);

// A batch whose (query tile x corpus split) items cannot all be resident at once is scanned in query chunks that can:
// with every list of a query live at the same time the lists share one tight threshold (DenseParams::sthr) and the
// epilogue stops being the bottleneck (cfg3: 4096 queries as 2 x 2048 run at ~1.3 PFLOP/s instead of ~1.2).  Each
// chunk streams the corpus again, which costs nothing while the scan is compute-bound (cfg4 shard, 12.5M x 1024-d,
// 8192 queries as 4 x 2048: 167.8 -> 159.8 ms; the re-reads are 102 GB per step = 0.64 TB/s).
template <int EPL>
static int run_dense_rows(const arb_index* ix, int64_t n_rows, const uint32_t* thr_init, const __nv_bfloat16* qplanes,
                          int q_planes, int64_t nq, int k, int64_t id_base, float* d_scores, int64_t* d_ids,
                          cudaStream_t st) {
  const int sms = num_sms();
  const int64_t n_qtiles = (nq + BLOCK_M - 1) / BLOCK_M;
  // the most query tiles one chunk may hold: items = qtiles * floor(sms / qtiles) splits all resident, and the lists of a
  // query (2 per split) must reach k with at most 8 keys each
  int64_t max_qt = 0;
  for (int64_t qt = 1; qt <= sms; ++qt) {
    const int64_t splits = sms / qt;
    if (splits >= 1 && (k + 2 * splits - 1) / (2 * splits) <= 8) max_qt = qt;
  }
  bool off = false;                                               // ARB_DENSE_CHUNK=0: A/B switch for the probes
  if (const char* v = std::getenv("ARB_DENSE_CHUNK")) off = atoi(v) == 0;
  const bool chunked = !off && max_qt > 0 && n_qtiles > max_qt;
  if (!chunked) return run_dense_chunk<EPL>(ix, n_rows, thr_init, qplanes, q_planes, nq, k, id_base, d_scores, d_ids, st);
  const int64_t n_chunks = (n_qtiles + max_qt - 1) / max_qt;
  const int64_t per = ((n_qtiles + n_chunks - 1) / n_chunks) * BLOCK_M;       // balanced, whole query tiles
  for (int64_t q0 = 0; q0 < nq; q0 += per) {
    const int64_t m = std::min<int64_t>(per, nq - q0);
    int rc = run_dense_chunk<EPL>(ix, n_rows, thr_init ? thr_init + q0 : nullptr,
                                  qplanes + static_cast<size_t>(q0) * q_planes * ix->d_pad, q_planes, m, k, id_base,
                                  d_scores + static_cast<size_t>(q0) * k, d_ids + static_cast<size_t>(q0) * k, st);
    if (rc) return rc;
  }
  return ARB_OK;
}

template <int EPL>
static int run_dense_chunk(const arb_index* ix, int64_t n_rows, const uint32_t* thr_init, const __nv_bfloat16* qplanes,
                           int q_planes, int64_t nq, int k, int64_t id_base, float* d_scores, int64_t* d_ids,
                           cudaStream_t st) {
  constexpr int C = 32 * EPL;
  DenseParams p;
  memset(&p, 0, sizeof(p));
  p.nq = static_cast<int>(nq);
  p.n_rows = static_cast<uint32_t>(n_rows);
  p.k = k;
  p.kb_per_term = ix->d_pad / BLOCK_K;
  setup_terms(p, q_planes, ix->planes, ix->d_pad);
  p.n_qtiles = static_cast<int>((nq + BLOCK_M - 1) / BLOCK_M);
  p.n_ntiles = static_cast<int>((n_rows + BLOCK_N - 1) / BLOCK_N);
  const int sms = num_sms();
  // CTA pairs (cta_group::2) for batches of at least 256 queries
  // (measured slower than the single-CTA kernel on cfg3 -- 6.2 vs 5.9 ms -- so it is opt-in: ARB_DENSE_PAIR=1)
  bool use_pair = false;
  if (const char* v = std::getenv("ARB_DENSE_PAIR")) use_pair = atoi(v) != 0 && nq >= 2 * BLOCK_M && sms >= 2;
  int grid, rounds = 0;
  if (use_pair) {
    p.n_qtiles = (p.n_qtiles + 1) & ~1;                  // whole 256-query super tiles (the odd one is padding)
    const int clusters = sms / 2;
    plan_splits(p.n_qtiles / 2, p.n_ntiles, clusters, &p.n_splits, &p.tiles_per_split);
    p.n_items = p.n_splits * (p.n_qtiles / 2);
    grid = 2 * std::min(clusters, p.n_items);
  } else {
    plan_splits(p.n_qtiles, p.n_ntiles, sms, &p.n_splits, &p.tiles_per_split, &rounds);
    p.n_items = p.n_splits * p.n_qtiles;
    grid = std::min(sms, p.n_items);
  }
  const size_t n_slots = static_cast<size_t>(p.n_splits) * p.n_qtiles;   // candidate lists per query row

  AsyncBuf lists, counts, gthr, sthr;
  bool cross_off = false;                                         // ARB_DENSE_CROSS=0: A/B switch for the probes
  if (const char* v = std::getenv("ARB_DENSE_CROSS")) cross_off = atoi(v) == 0;
  p.n_lists = p.n_splits * EPI_HALVES;
  p.rank2 = (k + p.n_lists - 1) / p.n_lists;
  if (!use_pair && !cross_off && rounds == 1 && p.n_items <= grid && p.n_lists > 1 && p.rank2 <= 8) {
    const size_t bytes = static_cast<size_t>(p.n_qtiles) * BLOCK_M * p.n_lists * sizeof(uint32_t);
    ARB_CUDA(sthr.alloc(bytes, st));
    ARB_CUDA(cudaMemsetAsync(sthr.p, 0, bytes, st));
    p.sthr = sthr.as<uint32_t>();
  }
  ARB_CUDA(lists.alloc(n_slots * BLOCK_M * EPI_HALVES * C * sizeof(uint2), st));
  ARB_CUDA(counts.alloc(n_slots * BLOCK_M * EPI_HALVES * sizeof(int), st));
  ARB_CUDA(gthr.alloc(static_cast<size_t>(p.n_qtiles) * BLOCK_M * sizeof(uint32_t), st));
  p.lists = lists.as<uint2>();
  p.counts = counts.as<int>();
  p.gthr = gthr.as<uint32_t>();
  const size_t ng = static_cast<size_t>(p.n_qtiles) * BLOCK_M;
  init_gthr_kernel<<<static_cast<unsigned>((ng + 255) / 256), 256, 0, st>>>(p.gthr, thr_init, p.nq, ng);
  ARB_LAUNCH_OK();

  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, qplanes, static_cast<uint64_t>(nq), static_cast<uint64_t>(q_planes) * ix->d_pad, BLOCK_M);
  if (rc) return rc;
  rc = make_tmap(&tb, ix->rows, static_cast<uint64_t>(n_rows), static_cast<uint64_t>(ix->planes) * ix->d_pad,
                 use_pair ? BLOCK_N / 2 : BLOCK_N);              // a pair's CTA stages half of the corpus tile
  if (rc) return rc;

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(dense_topk_kernel<EPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DENSE_SMEM);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(dense_topk_kernel<EPL, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, DENSE_SMEM);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(dense_topk_kernel<EPL, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, DENSE_SMEM);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(dense_topk_pair_kernel<EPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(dense_merge_kernel<EPL, false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(16 * 32 * EPL * sizeof(uint2)));
  });
  ARB_CUDA(attr_err);
  {
    if (use_pair) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(static_cast<unsigned>(grid));
      cfg.blockDim = dim3(DENSE_THREADS);
      cfg.dynamicSmemBytes = PAIR_SMEM;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      ARB_CUDA(cudaLaunchKernelEx(&cfg, dense_topk_pair_kernel<EPL>, ta, tb, p));
    } else {
      if (p.sthr && p.rank2 <= 4) dense_topk_kernel<EPL, false, 4><<<grid, DENSE_THREADS, DENSE_SMEM, st>>>(ta, tb, p);
      else if (p.sthr)            dense_topk_kernel<EPL, false, 8><<<grid, DENSE_THREADS, DENSE_SMEM, st>>>(ta, tb, p);
      else                        dense_topk_kernel<EPL, false><<<grid, DENSE_THREADS, DENSE_SMEM, st>>>(ta, tb, p);
    }
  }
  ARB_LAUNCH_OK();
  const int wpb = 4;
  {
    if (p.n_splits * EPI_HALVES >= 48 && nq <= 2048)
      dense_merge_kernel<EPL, false, 16><<<static_cast<unsigned>(nq), 16 * 32, 16 * C * sizeof(uint2), st>>>(p, id_base, d_scores, d_ids);
    else
      dense_merge_kernel<EPL, false, 4><<<static_cast<unsigned>(nq), 4 * 32, 4 * C * sizeof(uint2), st>>>(p, id_base, d_scores, d_ids);
  }
  ARB_LAUNCH_OK();
  return ARB_OK;
}

// Exact flat search.  Optionally the corpus is scanned twice: a pre-pass over the first tenth of the rows gives every
// query the k-th score of that sample, and the full scan starts from that bound instead of -inf.  The epilogue
// pays ~45 warp instructions whenever any of a warp's 32 queries has a score above its threshold in a group of 8
// columns; with cold thresholds that is ~17 % of all column steps and the four epilogue warps -- not the tensor
// pipe -- bound the kernel (6.7 ms vs 4.1 ms for the bare GEMM at cfg3).  The pre-pass costs a tenth of a scan.
template <int EPL>
static int run_dense(const arb_index* ix, const __nv_bfloat16* qplanes, int q_planes, int64_t nq, int k,
                     int64_t id_base, float* d_scores, int64_t* d_ids, cudaStream_t st) {
  // (with two epilogue warps per TMEM quadrant the pre-pass no longer pays on cfg3 -- 5.68 vs 5.69 ms -- so it is
  //  opt-in: ARB_DENSE_PREPASS=1; it was worth 1 ms with four epilogue warps)
  bool prepass = false;
  if (const char* v = std::getenv("ARB_DENSE_PREPASS")) prepass = atoi(v) != 0 && ix->n >= 262144 && nq >= BLOCK_M && k <= ix->n / 16;
  ProfileScope ps(PROF_DENSE_TOPK, st);
  if (!prepass) return run_dense_rows<EPL>(ix, ix->n, nullptr, qplanes, q_planes, nq, k, id_base, d_scores, d_ids, st);
  const int64_t n_pre = ((ix->n / 10 + BLOCK_N - 1) / BLOCK_N) * BLOCK_N;
  AsyncBuf thr;
  ARB_CUDA(thr.alloc(static_cast<size_t>(nq) * 4, st));
  int rc = run_dense_rows<EPL>(ix, n_pre, nullptr, qplanes, q_planes, nq, k, id_base, d_scores, d_ids, st);
  if (rc) return rc;
  prepass_bound_kernel<<<static_cast<unsigned>((nq + 255) / 256), 256, 0, st>>>(d_scores, d_ids, static_cast<int>(nq), k, thr.as<uint32_t>());
  ARB_LAUNCH_OK();
  return run_dense_rows<EPL>(ix, ix->n, thr.as<uint32_t>(), qplanes, q_planes, nq, k, id_base, d_scores, d_ids, st);
}

extern "C" int arb_index_search(const arb_index* ix, const void* queries, int dtype, int64_t nq, int queries_is_device,
                                int k, int query_planes, int64_t id_base, float* out_scores, int64_t* out_ids,
                                int out_is_device, void* stream) {
  ARB_CHECK(ix != nullptr, ARB_EINVAL, "arb_index_search: NULL index");
  ARB_CHECK(nq >= 0 && (queries != nullptr || nq == 0), ARB_EINVAL, "arb_index_search: bad queries/nq");
  ARB_CHECK(out_scores != nullptr && out_ids != nullptr, ARB_EINVAL, "arb_index_search: NULL output");
  ARB_CHECK(dtype == ARB_F32 || dtype == ARB_BF16, ARB_EINVAL, "arb_index_search: dtype must be ARB_F32 or ARB_BF16");
  ARB_CHECK(k >= 1 && k <= ARB_MAX_TOPK, ARB_ELIMIT, "arb_index_search: k must be in [1, 256]");
  ARB_CHECK(query_planes == 1 || query_planes == 3, ARB_EINVAL, "arb_index_search: query_planes must be 1 or 3");
  ARB_CHECK(nq <= (1 << 24), ARB_ELIMIT, "arb_index_search: at most 2^24 queries per call");
  ARB_CHECK(ix->n > 0, ARB_EINVAL, "arb_index_search: index is empty");
  if (nq == 0) return ARB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t esz = dtype == ARB_F32 ? 4 : 2;

  AsyncBuf qstage, qplanes, oscores, oids;
  const void* q_dev = queries;
  if (!queries_is_device) {
    ARB_CUDA(qstage.alloc(static_cast<size_t>(nq) * ix->dim * esz, st));
    ARB_CUDA(cudaMemcpyAsync(qstage.p, queries, static_cast<size_t>(nq) * ix->dim * esz, cudaMemcpyHostToDevice, st));
    q_dev = qstage.p;
  }
  ARB_CUDA(qplanes.alloc(static_cast<size_t>(nq) * query_planes * ix->d_pad * 2, st));
  int rc = launch_prep(q_dev, dtype, nq, ix->dim, ix->d_pad, query_planes, ix->normalize, qplanes.as<__nv_bfloat16>(), st);
  if (rc) return rc;

  float* d_scores = out_scores;
  int64_t* d_ids = out_ids;
  if (!out_is_device) {
    ARB_CUDA(oscores.alloc(static_cast<size_t>(nq) * k * 4, st));
    ARB_CUDA(oids.alloc(static_cast<size_t>(nq) * k * 8, st));
    d_scores = oscores.as<float>();
    d_ids = oids.as<int64_t>();
  }
  // candidate lists of 256 / 512 / 1024 entries: a smaller list is compacted more often, so its threshold stays
  // closer to the running k-th score and fewer columns take the append path
  if (k <= 100)      rc = run_dense<8>(ix, qplanes.as<__nv_bfloat16>(), query_planes, nq, k, id_base, d_scores, d_ids, st);
  else if (k <= 128) rc = run_dense<16>(ix, qplanes.as<__nv_bfloat16>(), query_planes, nq, k, id_base, d_scores, d_ids, st);
  else               rc = run_dense<32>(ix, qplanes.as<__nv_bfloat16>(), query_planes, nq, k, id_base, d_scores, d_ids, st);
  if (rc) return rc;
  if (!out_is_device) {
    ARB_CUDA(cudaMemcpyAsync(out_scores, d_scores, static_cast<size_t>(nq) * k * 4, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaMemcpyAsync(out_ids, d_ids, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaStreamSynchronize(st));
  }
  return ARB_OK;
}

// ============================================================================
// IVF (replaces faiss.IndexIVFFlat(IndexFlatIP(d), d, nlist, METRIC_INNER_PRODUCT) behind
// VectorIndex: vector_index.py:85-91 create, :126-169 train, :187-193 add_with_ids, :223 search).
//
// Build: the coarse quantiser is the exact top-1 inner-product search over the centroids (the same
// tcgen05 kernel with k = 1); rows are regrouped list-major on the device (stable radix sort by
// list, so a list keeps ascending ids) the first time a search follows an add.
// Search: top-nprobe centroids per query (same kernel, k = nprobe); the (query, probed list) pairs
// are bucketed by list on the device and the queries of each list are gathered into 128-row A
// tiles; one work item of the GEMM+select kernel = one A tile x the rows of its list, so a list
// is streamed once for all the queries probing it and the score block never reaches HBM.
// ============================================================================
namespace arb {

__global__ void ivf_ids_to_assign_kernel(const int64_t* __restrict__ ids, int64_t n, int32_t* __restrict__ assign) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) assign[i] = static_cast<int32_t>(ids[i] < 0 ? 0 : ids[i]);
}
__global__ void ivf_iota_kernel(int32_t* __restrict__ v, int64_t n) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) v[i] = static_cast<int32_t>(i);
}
// list_off[l] = first sorted position whose list >= l  (one thread per list boundary)
__global__ void ivf_listoff_kernel(const int32_t* __restrict__ sorted_list, int64_t n, int nlist, uint32_t* __restrict__ list_off) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l > nlist) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sorted_list[mid] < l) lo = mid + 1; else hi = mid; }
  list_off[l] = static_cast<uint32_t>(lo);
}
// regroup rows list-major: one warp per destination row
__global__ void __launch_bounds__(256)
ivf_regroup_kernel(const __nv_bfloat16* __restrict__ rows, const int32_t* __restrict__ perm, int64_t n, int row_elems,
                   __nv_bfloat16* __restrict__ g_rows, int64_t* __restrict__ g_ids) {
  const int lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const int64_t src = perm[r];
  if (lane == 0) g_ids[r] = src;
  const uint4* s4 = reinterpret_cast<const uint4*>(rows + src * row_elems);
  uint4* d4 = reinterpret_cast<uint4*>(g_rows + r * row_elems);
  for (int i = lane; i < row_elems / 8; i += 32) d4[i] = s4[i];
}

// ---- search-time planning -------------------------------------------------
__global__ void ivf_count_kernel(const int64_t* __restrict__ probe, int64_t n_pairs, int* __restrict__ cnt, int* __restrict__ slot) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n_pairs) return;
  const int64_t l = probe[i];
  slot[i] = l >= 0 ? atomicAdd(&cnt[l], 1) : -1;
}
// one block: items per list -> exclusive scan -> item table
__global__ void __launch_bounds__(1024)
ivf_plan_kernel(const int* __restrict__ cnt, int nlist, const uint32_t* __restrict__ list_off, int* __restrict__ item_start,
                int* __restrict__ n_items, IvfItem* __restrict__ items) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nlist; base += 1024) {
    const int l = base + tid;
    const int mine = l < nlist ? (cnt[l] + BLOCK_M - 1) / BLOCK_M : 0;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += v; }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const int carry = carry_s;
    const int start = carry + (warp ? warp_sum[warp - 1] : 0) + incl - mine;
    if (l < nlist) {
      item_start[l] = start;
      const IvfItem it = {list_off[l], list_off[l + 1]};
      for (int i = 0; i < mine; ++i) items[start + i] = it;
    }
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sum[31];
    __syncthreads();
  }
  if (tid == 0) *n_items = carry_s;
}
__global__ void ivf_place_kernel(const int64_t* __restrict__ probe, const int* __restrict__ slot, const int* __restrict__ item_start,
                                 int64_t n_pairs, int nprobe, int* __restrict__ row_query, int* __restrict__ pair_pos) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n_pairs) return;
  const int64_t l = probe[i];
  int pos = -1;
  if (l >= 0) {
    pos = item_start[l] * BLOCK_M + slot[i];
    row_query[pos] = static_cast<int>(i / nprobe);
  }
  pair_pos[i] = pos;
}
// gathered A matrix: row pos <- query planes of row_query[pos] (zeros for padding rows); one warp per row
__global__ void __launch_bounds__(256)
ivf_gather_kernel(const __nv_bfloat16* __restrict__ qplanes, const int* __restrict__ row_query, const int* __restrict__ n_items,
                  int row_elems, __nv_bfloat16* __restrict__ ag) {
  const int lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= static_cast<int64_t>(*n_items) * BLOCK_M) return;
  const int q = row_query[r];
  uint4* d4 = reinterpret_cast<uint4*>(ag + r * row_elems);
  if (q < 0) { for (int i = lane; i < row_elems / 8; i += 32) d4[i] = make_uint4(0u, 0u, 0u, 0u); return; }
  const uint4* s4 = reinterpret_cast<const uint4*>(qplanes + static_cast<int64_t>(q) * row_elems);
  for (int i = lane; i < row_elems / 8; i += 32) d4[i] = s4[i];
}

// ---- k-means (train_index) --------------------------------------------------
// new centroid = mean of the rows assigned to it (rows are visited in sorted order: deterministic);
// spherical: L2-normalised afterwards (faiss sets cp.spherical for inner-product IVF).  Block per centroid.
__global__ void __launch_bounds__(256)
ivf_kmeans_update_kernel(const __nv_bfloat16* __restrict__ rows, const int32_t* __restrict__ perm,
                         const uint32_t* __restrict__ list_off, int dim, int d_pad, int planes, int spherical,
                         float* __restrict__ centroids) {
  __shared__ float red[8];
  const int c = blockIdx.x;
  const uint32_t lo = list_off[c], hi = list_off[c + 1];
  if (lo == hi) return;                                   // empty cluster keeps its centroid
  const float inv = 1.0f / static_cast<float>(hi - lo);
  float ss = 0.f;
  for (int j = threadIdx.x; j < dim; j += blockDim.x) {
    float acc = 0.f;
    for (uint32_t i = lo; i < hi; ++i) {
      const __nv_bfloat16* src = rows + static_cast<int64_t>(perm[i]) * planes * d_pad;
      float v = 0.f;
      for (int pl = planes - 1; pl >= 0; --pl) v = __fadd_rn(v, __bfloat162float(src[pl * d_pad + j]));
      acc = __fadd_rn(acc, v);
    }
    acc = __fmul_rn(acc, inv);
    centroids[static_cast<int64_t>(c) * dim + j] = acc;
    ss = fmaf(acc, acc, ss);
  }
  if (!spherical) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float nrm = sqrtf(tot);
  if (nrm > 0.f)
    for (int j = threadIdx.x; j < dim; j += blockDim.x)
      centroids[static_cast<int64_t>(c) * dim + j] = __fdiv_rn(centroids[static_cast<int64_t>(c) * dim + j], nrm);
}
__global__ void ivf_pick_rows_kernel(const __nv_bfloat16* __restrict__ rows, int64_t n, int nlist, int dim, int d_pad, int planes,
                                     float* __restrict__ centroids) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<int64_t>(nlist) * dim) return;
  const int c = static_cast<int>(i / dim), j = static_cast<int>(i % dim);
  const int64_t r = static_cast<int64_t>(c) * n / nlist;                 // evenly spaced training rows as seeds
  const __nv_bfloat16* src = rows + r * planes * d_pad;
  float v = 0.f;
  for (int pl = planes - 1; pl >= 0; --pl) v = __fadd_rn(v, __bfloat162float(src[pl * d_pad + j]));
  centroids[i] = v;
}

}  // namespace arb

// rows in storage layout -> nearest centroid (max inner product), written to assign_out[0..n)
static int ivf_quantise(const arb_index* ix, const __nv_bfloat16* rows_storage, int64_t n, int32_t* assign_out, cudaStream_t st) {
  AsyncBuf sc, id;
  const int64_t chunk = int64_t(1) << 18;             // candidate lists are 4 KB per query row
  ARB_CUDA(sc.alloc(static_cast<size_t>(std::min(n, chunk)) * 4, st));
  ARB_CUDA(id.alloc(static_cast<size_t>(std::min(n, chunk)) * 8, st));
  for (int64_t off = 0; off < n; off += chunk) {
    const int64_t m = std::min(chunk, n - off);
    int rc = run_dense<16>(ix->coarse, rows_storage + static_cast<size_t>(off) * ix->row_elems(), ix->planes, m, 1, 0,
                           sc.as<float>(), id.as<int64_t>(), st);
    if (rc) return rc;
    ivf_ids_to_assign_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, st>>>(id.as<int64_t>(), m, assign_out + off);
    ARB_LAUNCH_OK();
  }
  return ARB_OK;
}

static int ivf_assign_rows(arb_index* ix, int64_t row0, int64_t n, cudaStream_t st) {
  return ivf_quantise(ix, ix->rows + static_cast<size_t>(row0) * ix->row_elems(), n, ix->assign + row0, st);
}

// (list, add-order position) stable sort -> perm (sorted positions) and list offsets
static int ivf_sort_by_list(const int32_t* assign, int64_t n, int nlist, AsyncBuf& perm, uint32_t* list_off, cudaStream_t st) {
  AsyncBuf keys_out, idx_in, tmp;
  ARB_CUDA(keys_out.alloc(static_cast<size_t>(n) * 4, st));
  ARB_CUDA(idx_in.alloc(static_cast<size_t>(n) * 4, st));
  ARB_CUDA(perm.alloc(static_cast<size_t>(n) * 4, st));
  ivf_iota_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(idx_in.as<int32_t>(), n);
  ARB_LAUNCH_OK();
  size_t tmp_bytes = 0;
  int end_bit = 1;
  while ((1 << end_bit) < nlist && end_bit < 31) ++end_bit;
  ARB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, assign, keys_out.as<int32_t>(), idx_in.as<int32_t>(),
                                           perm.as<int32_t>(), static_cast<int>(n), 0, end_bit, st));
  ARB_CUDA(tmp.alloc(tmp_bytes, st));
  ARB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, assign, keys_out.as<int32_t>(), idx_in.as<int32_t>(),
                                           perm.as<int32_t>(), static_cast<int>(n), 0, end_bit, st));
  count_launch(4);
  ivf_listoff_kernel<<<static_cast<unsigned>((nlist + 1 + 255) / 256), 256, 0, st>>>(keys_out.as<int32_t>(), n, nlist, list_off);
  ARB_LAUNCH_OK();
  return ARB_OK;
}

static int ivf_regroup(arb_index* ix, cudaStream_t st) {
  std::lock_guard<std::mutex> lock(ix->mu.m);                      // two first searches after an add must not both regroup
  if (!ix->grouped_dirty) return ARB_OK;
  ARB_CHECK(ix->n < (int64_t(1) << 31), ARB_ELIMIT, "IVF: at most 2^31 rows per shard");
  if (ix->g_cap < ix->n) {
    if (ix->g_rows) cudaFree(ix->g_rows);
    if (ix->g_ids) cudaFree(ix->g_ids);
    ix->g_rows = nullptr; ix->g_ids = nullptr; ix->g_cap = 0;
    if (cudaMalloc(&ix->g_rows, static_cast<size_t>(ix->n) * ix->row_elems() * 2) != cudaSuccess ||
        cudaMalloc(&ix->g_ids, static_cast<size_t>(ix->n) * 8) != cudaSuccess) {
      cudaGetLastError(); set_error("IVF regroup: out of device memory"); return ARB_ENOMEM;
    }
    ix->g_cap = ix->n;
  }
  AsyncBuf perm;
  int rc = ivf_sort_by_list(ix->assign, ix->n, ix->nlist, perm, ix->list_off, st);
  if (rc) return rc;
  ivf_regroup_kernel<<<static_cast<unsigned>((ix->n + 7) / 8), 256, 0, st>>>(ix->rows, perm.as<int32_t>(), ix->n,
                                                                            static_cast<int>(ix->row_elems()), ix->g_rows, ix->g_ids);
  ARB_LAUNCH_OK();
  ARB_CUDA(cudaStreamSynchronize(st));
  ix->g_n = ix->n;
  ix->grouped_dirty = false;
  return ARB_OK;
}

// centroids (fp32, device) -> coarse index; re-quantise the rows already stored
static int ivf_install_centroids(arb_index* ix, int nlist, cudaStream_t st) {
  if (ix->coarse) { arb_index_destroy(ix->coarse); ix->coarse = nullptr; }
  int rc = arb_index_create(&ix->coarse, ix->dim, ix->storage, 0);
  if (rc) return rc;
  rc = arb_index_add(ix->coarse, ix->centroids, ARB_F32, nlist, 1, st);
  if (rc) return rc;
  if (ix->list_off) cudaFree(ix->list_off);
  ix->list_off = nullptr;
  ARB_CUDA(cudaMalloc(&ix->list_off, static_cast<size_t>(nlist + 1) * 4));
  ix->nlist = nlist;
  ix->grouped_dirty = true;
  if (ix->n > 0) return ivf_assign_rows(ix, 0, ix->n, st);
  return ARB_OK;
}

extern "C" int arb_index_ivf_set_centroids(arb_index* ix, int nlist, const float* centroids, int is_device, void* stream) {
  ARB_CHECK(ix != nullptr && centroids != nullptr, ARB_EINVAL, "arb_index_ivf_set_centroids: NULL argument");
  ARB_CHECK(nlist >= 1 && nlist <= (1 << 20), ARB_ELIMIT, "arb_index_ivf_set_centroids: nlist must be in [1, 2^20]");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ix->centroids) { cudaFree(ix->centroids); ix->centroids = nullptr; }
  ARB_CUDA(cudaMalloc(&ix->centroids, static_cast<size_t>(nlist) * ix->dim * 4));
  ARB_CUDA(cudaMemcpyAsync(ix->centroids, centroids, static_cast<size_t>(nlist) * ix->dim * 4,
                           is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  int rc = ivf_install_centroids(ix, nlist, st);
  if (rc) return rc;
  ARB_CUDA(cudaStreamSynchronize(st));
  return ARB_OK;
}

extern "C" int arb_index_ivf_train(arb_index* ix, int nlist, const void* rows, int dtype, int64_t n, int rows_is_device,
                                   int niter, int spherical, void* stream) {
  ARB_CHECK(ix != nullptr && rows != nullptr, ARB_EINVAL, "arb_index_ivf_train: NULL argument");
  ARB_CHECK(nlist >= 1 && nlist <= (1 << 20), ARB_ELIMIT, "arb_index_ivf_train: nlist must be in [1, 2^20]");
  ARB_CHECK(n >= nlist, ARB_EINVAL, "arb_index_ivf_train: fewer training rows than lists");
  ARB_CHECK(n < (int64_t(1) << 31), ARB_ELIMIT, "arb_index_ivf_train: at most 2^31 training rows");
  ARB_CHECK(dtype == ARB_F32 || dtype == ARB_BF16, ARB_EINVAL, "arb_index_ivf_train: dtype must be ARB_F32 or ARB_BF16");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the training rows, pre-processed exactly like added rows, in a scratch index
  arb_index* tr = nullptr;
  int rc = arb_index_create(&tr, ix->dim, ix->storage, ix->normalize);
  if (rc) return rc;
  struct Guard { arb_index* p; ~Guard() { arb_index_destroy(p); } } guard{tr};
  rc = arb_index_add(tr, rows, dtype, n, rows_is_device, st);
  if (rc) return rc;
  if (ix->centroids) { cudaFree(ix->centroids); ix->centroids = nullptr; }
  ARB_CUDA(cudaMalloc(&ix->centroids, static_cast<size_t>(nlist) * ix->dim * 4));
  const int64_t nc = static_cast<int64_t>(nlist) * ix->dim;
  ivf_pick_rows_kernel<<<static_cast<unsigned>((nc + 255) / 256), 256, 0, st>>>(tr->rows, n, nlist, ix->dim, ix->d_pad, ix->planes, ix->centroids);
  ARB_LAUNCH_OK();
  AsyncBuf assign, loff;
  ARB_CUDA(assign.alloc(static_cast<size_t>(n) * 4, st));
  ARB_CUDA(loff.alloc(static_cast<size_t>(nlist + 1) * 4, st));
  for (int it = 0; it < niter; ++it) {
    // quantise with the current centroids, then move every centroid to the mean of its rows
    arb_index* coarse = nullptr;
    rc = arb_index_create(&coarse, ix->dim, ix->storage, 0);
    if (rc) return rc;
    Guard cg{coarse};
    rc = arb_index_add(coarse, ix->centroids, ARB_F32, nlist, 1, st);
    if (rc) return rc;
    arb_index view = *tr;                      // borrow tr's rows with this iteration's quantiser
    view.coarse = coarse;
    rc = ivf_quantise(&view, tr->rows, n, assign.as<int32_t>(), st);
    view.coarse = nullptr; view.rows = nullptr;
    if (rc) return rc;
    AsyncBuf perm;
    rc = ivf_sort_by_list(assign.as<int32_t>(), n, nlist, perm, loff.as<uint32_t>(), st);
    if (rc) return rc;
    ivf_kmeans_update_kernel<<<nlist, 256, 0, st>>>(tr->rows, perm.as<int32_t>(), loff.as<uint32_t>(), ix->dim, ix->d_pad,
                                                     ix->planes, spherical, ix->centroids);
    ARB_LAUNCH_OK();
    ARB_CUDA(cudaStreamSynchronize(st));       // coarse dies at scope end
  }
  rc = ivf_install_centroids(ix, nlist, st);
  if (rc) return rc;
  ARB_CUDA(cudaStreamSynchronize(st));
  return ARB_OK;
}

extern "C" int arb_index_ivf_nlist(const arb_index* ix) { return ix ? ix->nlist : 0; }

extern "C" int arb_index_ivf_export(const arb_index* ix, float* centroids_out, int32_t* assign_out, void* stream) {
  ARB_CHECK(ix != nullptr && ix->nlist > 0, ARB_EINVAL, "arb_index_ivf_export: index has no IVF state");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (centroids_out) ARB_CUDA(cudaMemcpyAsync(centroids_out, ix->centroids, static_cast<size_t>(ix->nlist) * ix->dim * 4, cudaMemcpyDeviceToHost, st));
  if (assign_out && ix->n) ARB_CUDA(cudaMemcpyAsync(assign_out, ix->assign, static_cast<size_t>(ix->n) * 4, cudaMemcpyDeviceToHost, st));
  ARB_CUDA(cudaStreamSynchronize(st));
  return ARB_OK;
}

template <int EPL>
static int run_ivf(arb_index* ix, const __nv_bfloat16* qplanes, int q_planes, int64_t nq, int k, int nprobe,
                   int64_t id_base, float* d_scores, int64_t* d_ids, cudaStream_t st) {
  constexpr int C = 32 * EPL;
  const int sms = num_sms();
  const int64_t n_pairs = nq * nprobe;
  const int64_t max_items = n_pairs / BLOCK_M + std::min<int64_t>(ix->nlist, n_pairs) + 1;
  ARB_CHECK(max_items * BLOCK_M < (int64_t(1) << 31), ARB_ELIMIT, "arb_index_ivf_search: nq * nprobe too large for one call");
  AsyncBuf pscore, pids, cnt, slot, item_start, n_items, items, row_query, pair_pos, ag, lists, counts, gthr;
  ARB_CUDA(pscore.alloc(static_cast<size_t>(n_pairs) * 4, st));
  ARB_CUDA(pids.alloc(static_cast<size_t>(n_pairs) * 8, st));
  {
    ProfileScope ps(PROF_IVF_COARSE, st);
    int rc = nprobe <= 128 ? run_dense<16>(ix->coarse, qplanes, q_planes, nq, nprobe, 0, pscore.as<float>(), pids.as<int64_t>(), st)
                           : run_dense<32>(ix->coarse, qplanes, q_planes, nq, nprobe, 0, pscore.as<float>(), pids.as<int64_t>(), st);
    if (rc) return rc;
  }
  ARB_CUDA(cnt.alloc(static_cast<size_t>(ix->nlist) * 4, st));
  ARB_CUDA(slot.alloc(static_cast<size_t>(n_pairs) * 4, st));
  ARB_CUDA(item_start.alloc(static_cast<size_t>(ix->nlist) * 4, st));
  ARB_CUDA(n_items.alloc(4, st));
  ARB_CUDA(items.alloc(static_cast<size_t>(max_items) * sizeof(IvfItem), st));
  ARB_CUDA(row_query.alloc(static_cast<size_t>(max_items) * BLOCK_M * 4, st));
  ARB_CUDA(pair_pos.alloc(static_cast<size_t>(n_pairs) * 4, st));
  const size_t q_row = static_cast<size_t>(q_planes) * ix->d_pad;
  ARB_CUDA(ag.alloc(static_cast<size_t>(max_items) * BLOCK_M * q_row * 2, st));
  ARB_CUDA(cudaMemsetAsync(cnt.p, 0, static_cast<size_t>(ix->nlist) * 4, st));
  ARB_CUDA(cudaMemsetAsync(row_query.p, 0xFF, static_cast<size_t>(max_items) * BLOCK_M * 4, st));
  ivf_count_kernel<<<static_cast<unsigned>((n_pairs + 255) / 256), 256, 0, st>>>(pids.as<int64_t>(), n_pairs, cnt.as<int>(), slot.as<int>());
  ARB_LAUNCH_OK();
  ivf_plan_kernel<<<1, 1024, 0, st>>>(cnt.as<int>(), ix->nlist, ix->list_off, item_start.as<int>(), n_items.as<int>(), items.as<IvfItem>());
  ARB_LAUNCH_OK();
  ivf_place_kernel<<<static_cast<unsigned>((n_pairs + 255) / 256), 256, 0, st>>>(pids.as<int64_t>(), slot.as<int>(), item_start.as<int>(), n_pairs,
                                                                                  nprobe, row_query.as<int>(), pair_pos.as<int>());
  ARB_LAUNCH_OK();
  ivf_gather_kernel<<<static_cast<unsigned>((max_items * BLOCK_M + 7) / 8), 256, 0, st>>>(qplanes, row_query.as<int>(), n_items.as<int>(),
                                                                                           static_cast<int>(q_row), ag.as<__nv_bfloat16>());
  ARB_LAUNCH_OK();

  DenseParams p;
  memset(&p, 0, sizeof(p));
  p.nq = static_cast<int>(nq);
  p.n_rows = static_cast<uint32_t>(ix->g_n);
  p.k = k;
  p.kb_per_term = ix->d_pad / BLOCK_K;
  setup_terms(p, q_planes, ix->planes, ix->d_pad);
  p.n_qtiles = static_cast<int>((nq + BLOCK_M - 1) / BLOCK_M);
  p.items = items.as<IvfItem>();
  p.n_items_ptr = n_items.as<int>();
  p.row_query = row_query.as<int>();
  p.pair_pos = pair_pos.as<int>();
  p.nprobe = nprobe;
  p.row_ids = ix->g_ids;
  ARB_CUDA(lists.alloc(static_cast<size_t>(max_items) * BLOCK_M * EPI_HALVES * C * sizeof(uint2), st));
  ARB_CUDA(counts.alloc(static_cast<size_t>(max_items) * BLOCK_M * EPI_HALVES * sizeof(int), st));
  ARB_CUDA(gthr.alloc(static_cast<size_t>(p.n_qtiles) * BLOCK_M * sizeof(uint32_t), st));
  p.lists = lists.as<uint2>();
  p.counts = counts.as<int>();
  p.gthr = gthr.as<uint32_t>();
  const size_t ng = static_cast<size_t>(p.n_qtiles) * BLOCK_M;
  fill_u32_kernel<<<static_cast<unsigned>((ng + 255) / 256), 256, 0, st>>>(p.gthr, KEY_NEG_INF, ng);
  ARB_LAUNCH_OK();

  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, ag.p, static_cast<uint64_t>(max_items) * BLOCK_M, q_row, BLOCK_M);
  if (rc) return rc;
  rc = make_tmap(&tb, ix->g_rows, static_cast<uint64_t>(ix->g_n), static_cast<uint64_t>(ix->planes) * ix->d_pad, BLOCK_N);
  if (rc) return rc;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(dense_topk_kernel<EPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DENSE_SMEM);
  });
  ARB_CUDA(attr_err);
  const int grid = static_cast<int>(std::min<int64_t>(sms, max_items));
  {
    ProfileScope ps(PROF_DENSE_TOPK, st);
    dense_topk_kernel<EPL, true><<<grid, DENSE_THREADS, DENSE_SMEM, st>>>(ta, tb, p);
  }
  ARB_LAUNCH_OK();
  const int wpb = 4;
  {
    ProfileScope ps(PROF_DENSE_MERGE, st);
    dense_merge_kernel<EPL, true, 4><<<static_cast<unsigned>(nq), 4 * 32, 4 * C * sizeof(uint2), st>>>(
        p, id_base, d_scores, d_ids);
  }
  ARB_LAUNCH_OK();
  return ARB_OK;
}

extern "C" int arb_index_ivf_search(arb_index* ix, const void* queries, int dtype, int64_t nq, int queries_is_device,
                                    int k, int nprobe, int query_planes, int64_t id_base, float* out_scores,
                                    int64_t* out_ids, int out_is_device, void* stream) {
  ARB_CHECK(ix != nullptr, ARB_EINVAL, "arb_index_ivf_search: NULL index");
  ARB_CHECK(ix->nlist > 0, ARB_EINVAL, "arb_index_ivf_search: index is not trained (no centroids)");
  ARB_CHECK(nq >= 0 && (queries != nullptr || nq == 0), ARB_EINVAL, "arb_index_ivf_search: bad queries/nq");
  ARB_CHECK(out_scores != nullptr && out_ids != nullptr, ARB_EINVAL, "arb_index_ivf_search: NULL output");
  ARB_CHECK(dtype == ARB_F32 || dtype == ARB_BF16, ARB_EINVAL, "arb_index_ivf_search: dtype must be ARB_F32 or ARB_BF16");
  ARB_CHECK(k >= 1 && k <= ARB_MAX_TOPK, ARB_ELIMIT, "arb_index_ivf_search: k must be in [1, 256]");
  ARB_CHECK(nprobe >= 1, ARB_EINVAL, "arb_index_ivf_search: nprobe must be >= 1");
  ARB_CHECK(query_planes == 1 || query_planes == 3, ARB_EINVAL, "arb_index_ivf_search: query_planes must be 1 or 3");
  ARB_CHECK(ix->n > 0, ARB_EINVAL, "arb_index_ivf_search: index is empty");
  if (nq == 0) return ARB_OK;
  nprobe = std::min(nprobe, std::min(ix->nlist, ARB_MAX_TOPK));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = ivf_regroup(ix, st);
  if (rc) return rc;
  const size_t esz = dtype == ARB_F32 ? 4 : 2;
  AsyncBuf qstage, qplanes, oscores, oids;
  const void* q_dev = queries;
  if (!queries_is_device) {
    ARB_CUDA(qstage.alloc(static_cast<size_t>(nq) * ix->dim * esz, st));
    ARB_CUDA(cudaMemcpyAsync(qstage.p, queries, static_cast<size_t>(nq) * ix->dim * esz, cudaMemcpyHostToDevice, st));
    q_dev = qstage.p;
  }
  ARB_CUDA(qplanes.alloc(static_cast<size_t>(nq) * query_planes * ix->d_pad * 2, st));
  rc = launch_prep(q_dev, dtype, nq, ix->dim, ix->d_pad, query_planes, ix->normalize, qplanes.as<__nv_bfloat16>(), st);
  if (rc) return rc;
  float* d_scores = out_scores;
  int64_t* d_ids = out_ids;
  if (!out_is_device) {
    ARB_CUDA(oscores.alloc(static_cast<size_t>(nq) * k * 4, st));
    ARB_CUDA(oids.alloc(static_cast<size_t>(nq) * k * 8, st));
    d_scores = oscores.as<float>();
    d_ids = oids.as<int64_t>();
  }
  if (k <= 100)      rc = run_ivf<8>(ix, qplanes.as<__nv_bfloat16>(), query_planes, nq, k, nprobe, id_base, d_scores, d_ids, st);
  else if (k <= 128) rc = run_ivf<16>(ix, qplanes.as<__nv_bfloat16>(), query_planes, nq, k, nprobe, id_base, d_scores, d_ids, st);
  else               rc = run_ivf<32>(ix, qplanes.as<__nv_bfloat16>(), query_planes, nq, k, nprobe, id_base, d_scores, d_ids, st);
  if (rc) return rc;
  if (!out_is_device) {
    ARB_CUDA(cudaMemcpyAsync(out_scores, d_scores, static_cast<size_t>(nq) * k * 4, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaMemcpyAsync(out_ids, d_ids, static_cast<size_t>(nq) * k * 8, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaStreamSynchronize(st));
  }
  return ARB_OK;
}

// ----------------------------------------------------------------------------
// Per-candidate similarities (replaces the Python loop of QueryProcessor._calculate_vector_similarities,
// query/query_processor.py:3543-3589: cosine of the query with the stored embedding of every note of a
// candidate pool, clamped at 0).  One warp per (query, candidate): both vectors are read back from their
// bf16 planes (hi + mid + lo = the stored fp32 value), fp32 FMA dot product, warp reduction.
// ----------------------------------------------------------------------------
namespace arb {
__global__ void __launch_bounds__(256)
score_rows_kernel(const __nv_bfloat16* __restrict__ rows, int64_t n_rows, int planes, const __nv_bfloat16* __restrict__ qplanes,
                  int q_planes, int d_pad, const int64_t* __restrict__ cand, int64_t nq, int nc, int clamp_nonneg,
                  float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t pair = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pair >= nq * nc) return;
  const int64_t q = pair / nc;
  const int64_t r = cand[pair];
  float s = 0.f;
  if (r >= 0 && r < n_rows) {
    const __nv_bfloat16* x = rows + r * static_cast<int64_t>(planes) * d_pad;
    const __nv_bfloat16* y = qplanes + q * static_cast<int64_t>(q_planes) * d_pad;
    for (int i = lane; i < d_pad; i += 32) {
      float a = 0.f, b = 0.f;
      for (int pl = planes - 1; pl >= 0; --pl) a = __fadd_rn(a, __bfloat162float(x[pl * d_pad + i]));
      for (int pl = q_planes - 1; pl >= 0; --pl) b = __fadd_rn(b, __bfloat162float(y[pl * d_pad + i]));
      s = fmaf(a, b, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (clamp_nonneg) s = fmaxf(s, 0.f);                          // max(0.0, similarity) (:3567)
  }
  if (lane == 0) out[pair] = s;
}
}  // namespace arb

extern "C" int arb_index_score_rows(const arb_index* ix, const void* queries, int dtype, int64_t nq, int queries_is_device,
                                    const int64_t* cand_rows, int nc, int cand_is_device, int query_planes,
                                    int clamp_nonneg, float* out, int out_is_device, void* stream) {
  ARB_CHECK(ix != nullptr, ARB_EINVAL, "arb_index_score_rows: NULL index");
  ARB_CHECK(nq >= 0 && nc >= 0 && (nq == 0 || nc == 0 || (queries && cand_rows && out)), ARB_EINVAL, "arb_index_score_rows: bad argument");
  ARB_CHECK(dtype == ARB_F32 || dtype == ARB_BF16, ARB_EINVAL, "arb_index_score_rows: dtype must be ARB_F32 or ARB_BF16");
  ARB_CHECK(query_planes == 1 || query_planes == 3, ARB_EINVAL, "arb_index_score_rows: query_planes must be 1 or 3");
  if (nq == 0 || nc == 0) return ARB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t esz = dtype == ARB_F32 ? 4 : 2;
  const int64_t n_pairs = nq * nc;
  AsyncBuf qstage, qplanes, cstage, obuf;
  const void* q_dev = queries;
  if (!queries_is_device) {
    ARB_CUDA(qstage.alloc(static_cast<size_t>(nq) * ix->dim * esz, st));
    ARB_CUDA(cudaMemcpyAsync(qstage.p, queries, static_cast<size_t>(nq) * ix->dim * esz, cudaMemcpyHostToDevice, st));
    q_dev = qstage.p;
  }
  ARB_CUDA(qplanes.alloc(static_cast<size_t>(nq) * query_planes * ix->d_pad * 2, st));
  int rc = launch_prep(q_dev, dtype, nq, ix->dim, ix->d_pad, query_planes, ix->normalize, qplanes.as<__nv_bfloat16>(), st);
  if (rc) return rc;
  const int64_t* d_cand = cand_rows;
  if (!cand_is_device) {
    ARB_CUDA(cstage.alloc(static_cast<size_t>(n_pairs) * 8, st));
    ARB_CUDA(cudaMemcpyAsync(cstage.p, cand_rows, static_cast<size_t>(n_pairs) * 8, cudaMemcpyHostToDevice, st));
    d_cand = cstage.as<int64_t>();
  }
  float* d_out = out;
  if (!out_is_device) { ARB_CUDA(obuf.alloc(static_cast<size_t>(n_pairs) * 4, st)); d_out = obuf.as<float>(); }
  score_rows_kernel<<<static_cast<unsigned>((n_pairs + 7) / 8), 256, 0, st>>>(ix->rows, ix->n, ix->planes, qplanes.as<__nv_bfloat16>(),
                                                                             query_planes, ix->d_pad, d_cand, nq, nc, clamp_nonneg, d_out);
  ARB_LAUNCH_OK();
  if (!out_is_device) {
    ARB_CUDA(cudaMemcpyAsync(out, d_out, static_cast<size_t>(n_pairs) * 4, cudaMemcpyDeviceToHost, st));
    ARB_CUDA(cudaStreamSynchronize(st));
  }
  return ARB_OK;
}
