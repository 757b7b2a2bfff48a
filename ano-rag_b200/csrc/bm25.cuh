// Declarations shared by the BM25 kernels (bm25.cu: index build, per-query kernel,
// full score vectors; bm25_batch.cu: the batched plane/sparse pipeline).
#pragma once

#include "common.cuh"

#include <vector>

namespace arb {

constexpr int DOC_BLOCK_LOG2 = 14;
constexpr int DOC_BLOCK = 1 << DOC_BLOCK_LOG2;       // docs per posting block (blk_off granularity)
constexpr int GCODE_BITS = 18;
constexpr uint32_t GCODE_MASK = (1u << GCODE_BITS) - 1u;
__host__ __device__ __forceinline__ uint32_t post_doc(uint32_t u) { return u & (DOC_BLOCK - 1); }
__host__ __device__ __forceinline__ uint32_t post_code(uint32_t u) { return u >> DOC_BLOCK_LOG2; }

constexpr int SLICE_LOG2 = DOC_BLOCK_LOG2 - 5;       // 512 docs: fine_off granularity
constexpr int SLICE = 1 << SLICE_LOG2;

// Doc-major "planes" of the densest terms (batched pipeline): plane s holds, for every doc,
// the term's whole contribution idf*g as fp64 (+0.0 where the term is absent), stored
// tile-major: planes[(doc >> 8) * n_planes + s][plane_pos(doc & 255)].
constexpr int PLANE_TILE_LOG2 = 8;
constexpr int PLANE_TILE = 1 << PLANE_TILE_LOG2;
constexpr int MAX_PLANES = 96;
// Position of doc d (0..255) inside a 256-entry plane-tile row.  The plane pass gives lane l the eight
// docs 8l .. 8l+7 and reads them as four 16-byte chunks; chunk j of all lanes is contiguous, so every
// LDS.128 of a warp is conflict free:  d = 8l + 2j + h  ->  (j*32 + l)*2 + h.
__host__ __device__ __forceinline__ uint32_t plane_pos(uint32_t d) {
  return ((((d >> 1) & 3u) * 32u + (d >> 3)) << 1) | (d & 1u);
}

struct Bm25Dev {
  int64_t n_docs, n_terms, n_post;
  int nb;                       // 16K-doc posting blocks
  int n_g;
  const uint32_t* blk_off;      // [n_terms*nb + 1]
  const uint32_t* post;         // [n_post]  gcode << 14 | doc % 16384
  const double* idf;            // [n_terms]
  const double* gtab;           // [n_g]
  int n_slices;                 // 512-doc slices
  int n_dense;                  // terms with per-slice offsets
  const int* dense_slot;        // [n_terms] index into fine_off or -1
  const uint32_t* fine_off;     // [n_dense * (n_slices + 1)]
  int n_planes;                 // terms stored as doc-major planes (<= MAX_PLANES)
  const int* plane_slot;        // [n_terms] plane index or -1
  const double* planes;         // [n_tiles * n_planes * 256]
  const double* plane_tmax;     // [n_tiles * n_planes] largest value of each plane inside each tile (>= +0.0)
  const double* plane_gmax;     // [n_planes] largest value of each plane
  int64_t n_tiles;              // 256-doc tiles
  // fp32 copy of the planes, every value rounded UP (screening pass: an fp32 sum with upward rounding bounds
  // the in-order fp64 sum from above), tile-major with pos32() inside a row
  const float* planes32;        // [n_tiles * n_planes * 256]
  // doc-major rows of the non-plane ("sparse") postings with idf != 0, sorted by term inside a doc; the value is
  // the posting's finished contribution idf * g (one fp64 rounding, as everywhere else)
  int64_t n_sp;
  const uint32_t* drow_off;     // [n_docs + 1]
  const uint32_t* drow_term;    // [n_sp]
  const double* drow_val;       // [n_sp]
  // the same sparse postings once more, tile-major and TERM-major inside a 256-doc tile (docs ascending inside a
  // term): the member pass applies one decoded query-term record to all docs of its term in the tile
  const uint32_t* tt_off;       // [n_tiles + 1]
  const uint32_t* tt_term;      // [n_sp]
  const uint8_t* tt_doc;        // [n_sp] doc inside the tile
  const uint32_t* tt_info;      // [n_sp] doc inside the tile | (postings from here to the end of the term's run - 1) << 8 | first of its run << 31
  const double* tt_val;         // [n_sp]
  const uint32_t* tt_bloom;     // [n_tiles * 256 * BLOOM_W] per-doc filter of the doc's sparse terms (bloom_word / bloom_mask)
};

// Per-doc filter of the doc's sparse terms: 16 words; a term sets three of bits 0..30 of ONE word, so a look-up is one
// load and one masked compare (~1 % false positives per look-up for a doc with 22 sparse terms).  Bit 31 is never set:
// a record entry without a sparse term tests for it.
constexpr int BLOOM_W = 16;
__host__ __device__ __forceinline__ uint32_t bloom_word(uint32_t t) { return (t * 0x9E3779B1u) >> 28; }
__host__ __device__ __forceinline__ uint32_t bloom_mask(uint32_t t) {
  const uint32_t h = t * 0x85EBCA6Bu;
  return (1u << ((h >> 27) * 31u >> 5)) | (1u << (((h >> 22) & 31u) * 31u >> 5)) | (1u << (((h >> 17) & 31u) * 31u >> 5));
}

// Position of doc d (0..255) inside a 256-entry fp32 plane-tile row: lane l owns docs 8l .. 8l+7 and reads them as
// two 16-byte chunks; chunk j of all lanes is contiguous (conflict-free LDS.128):  d = 8l + 4j + e -> (j*32 + l)*4 + e.
__host__ __device__ __forceinline__ uint32_t pos32(uint32_t d) {
  return ((((d >> 2) & 1u) * 32u + (d >> 3)) << 2) | (d & 3u);
}

int merge_topk_device(int score_dtype, const void* scores, const int64_t* ids, int parts, int64_t nq, int k,
                      void* out_scores, int64_t* out_ids, cudaStream_t st);   // fuse.cu

// per-query kernel (bm25.cu) over the queries listed in flag_list[0 .. *flag_count) (NULL = all nq queries)
int launch_bm25_topk(const Bm25Dev& ix, const int32_t* d_terms, int max_qt, int k, int64_t id_base, int64_t nq,
                     int n_split, int blocks_per_split, const int* flag_list, const int* flag_count, int grid_x,
                     double* part_scores, int64_t* part_ids, cudaStream_t st);
int bm25_set_smem_attrs();

// batched pipeline (bm25_batch.cu); returns ARB_OK after enqueueing everything on `st`
int bm25_batch_search(const Bm25Dev& ix, const int32_t* d_terms, int64_t nq, int max_qt, int k, int64_t id_base,
                      double* d_scores, int64_t* d_ids, cudaStream_t st);

}  // namespace arb

struct arb_bm25 {
  arb::Bm25Dev dev;
  uint32_t* blk_off = nullptr;
  uint32_t* post = nullptr;
  double* idf = nullptr;
  double* gtab = nullptr;
  int* dense_slot = nullptr;
  uint32_t* fine_off = nullptr;
  int* plane_slot = nullptr;
  double* planes = nullptr;
  double* plane_tmax = nullptr;   // followed by plane_gmax
  float* planes32 = nullptr;
  uint32_t* drow_off = nullptr;
  uint32_t* drow_term = nullptr;
  double* drow_val = nullptr;
  uint32_t* tt_off = nullptr;
  uint32_t* tt_term = nullptr;
  uint8_t* tt_doc = nullptr;
  uint32_t* tt_info = nullptr;
  double* tt_val = nullptr;
  uint32_t* tt_bloom = nullptr;
  std::vector<int64_t> df;      // host copy of per-term df (roofline accounting)
};
