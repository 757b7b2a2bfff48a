// TF-IDF cosine scoring: the "BM25 fallback" of the reference's VectorRetriever
// (vector_store/retriever.py:924-1002).  The reference fits sklearn's TfidfVectorizer over the note texts and
// answers a query with cosine_similarity(query_vector, tfidf_matrix) + argsort.  Text -> sparse rows stays on the
// host (the vectorizer is the reference's own dependency, like the tokenizer); the library holds the
// L2-normalised document matrix term-major in HBM and evaluates the sparse product and the top-k.
//
// Bit-exactness: scipy's csr_matmat accumulates, for every document, v * B[term, doc] over the query's terms in
// ascending term order, each product and each sum rounded separately (x86-64 wheels carry no FMA).  The kernel does
// the same with __dmul_rn / __dadd_rn: one CTA per (doc block, query), the query's terms applied one after the
// other, the postings of a term spread over the threads (a doc occurs once per term, so there are no conflicts).
#include "common.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

namespace arb {

constexpr int TF_BLOCK_LOG2 = 12;
constexpr int TF_BLOCK = 1 << TF_BLOCK_LOG2;         // docs per block: 32 KB of fp64 accumulators in shared memory
constexpr int TF_THREADS = 256;

// scores of docs [b * TF_BLOCK, ...) for query q, written as sortable keys in REVERSE doc order (see tfidf_search):
// key_out[q * n_docs + (n_docs - 1 - doc)], idx_out[...] = doc
__global__ void __launch_bounds__(TF_THREADS)
tfidf_scores_kernel(int64_t n_docs, int nb, const int64_t* __restrict__ blk_off, const int32_t* __restrict__ post_doc,
                    const double* __restrict__ post_val, const int64_t* __restrict__ q_ptr,
                    const int32_t* __restrict__ q_term, const double* __restrict__ q_val,
                    unsigned long long* __restrict__ key_out, int32_t* __restrict__ idx_out) {
  __shared__ double acc[TF_BLOCK];
  const int b = blockIdx.x;
  const int64_t q = blockIdx.y;
  const int64_t d0 = static_cast<int64_t>(b) << TF_BLOCK_LOG2;
  for (int i = threadIdx.x; i < TF_BLOCK; i += TF_THREADS) acc[i] = 0.0;
  __syncthreads();
  for (int64_t i = q_ptr[q]; i < q_ptr[q + 1]; ++i) {          // ascending term id: scipy's accumulation order
    const int64_t t = q_term[i];
    const double w = q_val[i];
    const int64_t lo = blk_off[t * nb + b], hi = blk_off[t * nb + b + 1];
    for (int64_t p = lo + threadIdx.x; p < hi; p += TF_THREADS) {
      const int d = post_doc[p] - static_cast<int>(d0);
      acc[d] = __dadd_rn(acc[d], __dmul_rn(w, post_val[p]));
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < TF_BLOCK; i += TF_THREADS) {
    const int64_t d = d0 + i;
    if (d < n_docs) {
      const int64_t at = q * n_docs + (n_docs - 1 - d);
      key_out[at] = f64_key(acc[i]);
      idx_out[at] = static_cast<int32_t>(d);
    }
  }
}

__global__ void tfidf_offsets_kernel(int64_t* off, int64_t nq, int64_t n_docs) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i <= nq) off[i] = i * n_docs;
}

// first k entries of every sorted segment; `similarities[idx] > 0` only (retriever.py:985)
__global__ void tfidf_emit_kernel(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ idx, int64_t nq,
                                  int64_t n_docs, int k, double* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= nq * k) return;
  const int64_t q = i / k, j = i % k;
  double s = 0.0; int64_t id = -1;
  if (j < n_docs) {
    const double v = key_f64(keys[q * n_docs + j]);
    if (v > 0.0) { s = v; id = idx[q * n_docs + j]; }
  }
  out_scores[i] = s;
  out_ids[i] = id;
}


// ----------------------------------------------------------------------------
// top-k of one query's N similarities without sorting them: an MSD radix select over the order-preserving keys finds the
// k-th largest positive key T in eight 8-bit passes (256-bin shared-memory histograms), the entries above T plus the
// first ones equal to T in position order (np.argsort(...)[::-1] semantics: the larger doc id first among equals; the
// keys are stored in reverse doc order) are collected and only those <= 1024 entries are sorted.  One CTA per query.
// ----------------------------------------------------------------------------
constexpr int SEL_THREADS = 1024;
constexpr int SEL_MAX_K = 1024;
__global__ void __launch_bounds__(SEL_THREADS)
tfidf_select_kernel(const unsigned long long* __restrict__ keys, int64_t n_docs, int k, double* __restrict__ out_scores,
                    int64_t* __restrict__ out_ids) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long ckey[SEL_MAX_K];
  __shared__ unsigned int cpos[SEL_MAX_K];
  __shared__ unsigned long long s_prefix;
  __shared__ unsigned int s_need, s_ngt, s_base, s_wsum[SEL_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t q = blockIdx.x;
  const unsigned long long* K = keys + q * n_docs;
  const unsigned long long K0 = f64_key(0.0);                  // only similarities > 0 are results (retriever.py:985)
  unsigned long long prefix = 0ull;
  unsigned int need = static_cast<unsigned int>(k);
  bool take_all = false;
  for (int pass = 0; pass < 8 && !take_all; ++pass) {
    const int shift = 56 - 8 * pass;
    if (tid < 256) hist[tid] = 0u;
    __syncthreads();
    for (int64_t i = tid; i < n_docs; i += SEL_THREADS) {
      const unsigned long long key = K[i];
      if (key > K0 && (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))))
        atomicAdd(&hist[static_cast<unsigned int>(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int above = 0u;
      int d = 255;
      for (; d >= 0; --d) {
        if (above + hist[d] >= need) break;
        above += hist[d];
      }
      if (d < 0) { s_need = 0xFFFFFFFFu; }                    // fewer positive entries than k (only possible in pass 0): all of them
      else { s_need = need - above; s_prefix = prefix | (static_cast<unsigned long long>(d) << shift); }
    }
    __syncthreads();
    if (s_need == 0xFFFFFFFFu) take_all = true;
    else { need = s_need; prefix = s_prefix; }
    __syncthreads();
  }
  // T = k-th largest positive key; `need` entries equal to T are wanted (the first in position order)
  const unsigned long long T = take_all ? K0 : prefix;
  if (take_all) need = 0u;
  if (tid == 0) { s_ngt = 0u; s_base = 0u; }
  __syncthreads();
  const unsigned int n_gt_expected = take_all ? 0xFFFFFFFFu : static_cast<unsigned int>(k) - need;
  for (int64_t c0 = 0; c0 < n_docs; c0 += SEL_THREADS) {
    const int64_t i = c0 + tid;
    const unsigned long long key = i < n_docs ? K[i] : 0ull;
    const bool gt = i < n_docs && key > T;
    const bool eq = !take_all && i < n_docs && key == T;
    if (gt) {
      const unsigned int at = atomicAdd(&s_ngt, 1u);
      if (at < static_cast<unsigned int>(SEL_MAX_K)) { ckey[at] = key; cpos[at] = static_cast<unsigned int>(i); }
    }
    // ordered ranks of the entries equal to T
    const unsigned int bal = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) s_wsum[warp] = __popc(bal);
    __syncthreads();
    unsigned int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_wsum[w];
    if (eq) {
      const unsigned int r = before + __popc(bal & ((1u << lane) - 1u));
      if (r < need) { ckey[n_gt_expected + r] = key; cpos[n_gt_expected + r] = static_cast<unsigned int>(i); }
    }
    __syncthreads();
    if (tid == 0) { unsigned int t = 0u; for (int w = 0; w < SEL_THREADS / 32; ++w) t += s_wsum[w]; s_base += t; }
    __syncthreads();
  }
  const int cnt = take_all ? static_cast<int>(min(s_ngt, static_cast<unsigned int>(SEL_MAX_K))) : k;
  int n2 = 2;
  while (n2 < cnt) n2 <<= 1;
  for (int i = cnt + tid; i < n2; i += SEL_THREADS) { ckey[i] = 0ull; cpos[i] = 0xFFFFFFFFu; }
  __syncthreads();
  // bitonic sort by (key desc, position asc)
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < n2 / 2; i += SEL_THREADS) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool best_first = ((lo & size) == 0);
        const unsigned long long ka = ckey[lo], kb = ckey[hi];
        const unsigned int pa = cpos[lo], pb = cpos[hi];
        const bool a_before_b = (ka > kb) || (ka == kb && pa < pb);
        if (a_before_b != best_first) { ckey[lo] = kb; ckey[hi] = ka; cpos[lo] = pb; cpos[hi] = pa; }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < k; j += SEL_THREADS) {
    const bool ok = j < cnt;
    out_scores[q * k + j] = ok ? key_f64(ckey[j]) : 0.0;
    out_ids[q * k + j] = ok ? n_docs - 1 - static_cast<int64_t>(cpos[j]) : -1;
  }
}

}  // namespace arb

using namespace arb;

struct arb_tfidf {
  int64_t n_docs = 0, n_terms = 0, n_post = 0;
  int nb = 0;
  int64_t* blk_off = nullptr;
  int32_t* post_doc = nullptr;
  double* post_val = nullptr;
};

extern "C" int arb_tfidf_destroy(arb_tfidf* ix) {
  if (!ix) return ARB_OK;
  cudaFree(ix->blk_off); cudaFree(ix->post_doc); cudaFree(ix->post_val);
  delete ix;
  return ARB_OK;
}

extern "C" int arb_tfidf_create(arb_tfidf** out, int64_t n_docs, int64_t n_terms, const int64_t* term_ptr,
                                const int32_t* post_doc, const double* post_val) {
  ARB_CHECK(out && term_ptr, ARB_EINVAL, "arb_tfidf_create: NULL argument");
  ARB_CHECK(n_docs >= 1 && n_docs <= 2147483647ll && n_terms >= 1, ARB_EINVAL, "arb_tfidf_create: bad shape");
  const int64_t n_post = term_ptr[n_terms];
  ARB_CHECK(n_post == 0 || (post_doc && post_val), ARB_EINVAL, "arb_tfidf_create: NULL postings");
  arb_tfidf* ix = new arb_tfidf();
  ix->n_docs = n_docs; ix->n_terms = n_terms; ix->n_post = n_post;
  ix->nb = static_cast<int>((n_docs + TF_BLOCK - 1) >> TF_BLOCK_LOG2);
  // block offsets: postings of a term are sorted by doc, so one sweep per term finds where each doc block starts
  std::vector<int64_t> off(static_cast<size_t>(n_terms) * ix->nb + 1);
  for (int64_t t = 0; t < n_terms; ++t) {
    int64_t p = term_ptr[t];
    const int64_t end = term_ptr[t + 1];
    int32_t prev = -1;
    for (int b = 0; b < ix->nb; ++b) {
      off[static_cast<size_t>(t) * ix->nb + b] = p;
      const int64_t lim = static_cast<int64_t>(b + 1) << TF_BLOCK_LOG2;
      while (p < end && post_doc[p] < lim) {
        if (post_doc[p] <= prev || post_doc[p] >= n_docs) {
          delete ix;
          set_error("arb_tfidf_create: postings of a term must be strictly ascending doc ids below n_docs");
          return ARB_EINVAL;
        }
        prev = post_doc[p];
        ++p;
      }
    }
    if (p != end) { delete ix; set_error("arb_tfidf_create: doc id out of range"); return ARB_EINVAL; }
  }
  off[static_cast<size_t>(n_terms) * ix->nb] = n_post;
  cudaError_t e = cudaMalloc(&ix->blk_off, off.size() * 8);
  if (e == cudaSuccess) e = cudaMalloc(&ix->post_doc, std::max<int64_t>(1, n_post) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&ix->post_val, std::max<int64_t>(1, n_post) * 8);
  if (e == cudaSuccess) e = cudaMemcpy(ix->blk_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && n_post) e = cudaMemcpy(ix->post_doc, post_doc, n_post * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && n_post) e = cudaMemcpy(ix->post_val, post_val, n_post * 8, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { arb_tfidf_destroy(ix); return cuda_fail(e, "arb_tfidf_create", __FILE__, __LINE__); }
  *out = ix;
  return ARB_OK;
}

extern "C" int arb_tfidf_search(const arb_tfidf* ix, int64_t nq, const int64_t* q_ptr, const int32_t* q_term,
                                const double* q_val, int k, double* out_scores, int64_t* out_ids, void* stream) {
  ARB_CHECK(ix && q_ptr && out_scores && out_ids, ARB_EINVAL, "arb_tfidf_search: NULL argument");
  ARB_CHECK(nq >= 0 && k >= 1, ARB_EINVAL, "arb_tfidf_search: bad nq / k");
  if (nq == 0) return ARB_OK;
  const int64_t nnz = q_ptr[nq];
  ARB_CHECK(nnz == 0 || (q_term && q_val), ARB_EINVAL, "arb_tfidf_search: NULL query terms");
  for (int64_t q = 0; q < nq; ++q)
    for (int64_t i = q_ptr[q]; i < q_ptr[q + 1]; ++i) {
      ARB_CHECK(q_term[i] >= 0 && q_term[i] < ix->n_terms, ARB_EINVAL, "arb_tfidf_search: term id out of range");
      ARB_CHECK(i == q_ptr[q] || q_term[i] > q_term[i - 1], ARB_EINVAL, "arb_tfidf_search: query terms must be ascending");
    }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n = ix->n_docs;
  // a chunk of queries at a time: 2 x (8 + 4) bytes per (query, doc) for the sort's double buffers
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(nq, (int64_t{1} << 30) / (n * 24)));
  AsyncBuf d_ptr, d_term, d_val, key_a, key_b, idx_a, idx_b, seg, tmp, o_s, o_i;
  ARB_CUDA(d_ptr.alloc((nq + 1) * 8, st));
  ARB_CUDA(d_term.alloc(std::max<int64_t>(1, nnz) * 4, st));
  ARB_CUDA(d_val.alloc(std::max<int64_t>(1, nnz) * 8, st));
  ARB_CUDA(cudaMemcpyAsync(d_ptr.p, q_ptr, (nq + 1) * 8, cudaMemcpyHostToDevice, st));
  if (nnz) {
    ARB_CUDA(cudaMemcpyAsync(d_term.p, q_term, nnz * 4, cudaMemcpyHostToDevice, st));
    ARB_CUDA(cudaMemcpyAsync(d_val.p, q_val, nnz * 8, cudaMemcpyHostToDevice, st));
  }
  ARB_CUDA(key_a.alloc(chunk * n * 8, st)); ARB_CUDA(key_b.alloc(chunk * n * 8, st));
  ARB_CUDA(idx_a.alloc(chunk * n * 4, st)); ARB_CUDA(idx_b.alloc(chunk * n * 4, st));
  ARB_CUDA(seg.alloc((chunk + 1) * 8, st));
  ARB_CUDA(o_s.alloc(nq * k * 8, st)); ARB_CUDA(o_i.alloc(nq * k * 8, st));
  tfidf_offsets_kernel<<<static_cast<unsigned>((chunk + 256) / 256), 256, 0, st>>>(seg.as<int64_t>(), chunk, n);
  ARB_LAUNCH_OK();
  size_t tmp_bytes = 0;
  cub::DoubleBuffer<unsigned long long> dk(key_a.as<unsigned long long>(), key_b.as<unsigned long long>());
  cub::DoubleBuffer<int32_t> dv(idx_a.as<int32_t>(), idx_b.as<int32_t>());
  ARB_CUDA(cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, tmp_bytes, dk, dv, chunk * n, static_cast<int>(chunk),
                                                              seg.as<int64_t>(), seg.as<int64_t>() + 1, 0, 64, st));
  ARB_CUDA(tmp.alloc(tmp_bytes, st));
  for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
    const int64_t m = std::min(chunk, nq - q0);
    dk.selector = 0; dv.selector = 0;
    dim3 grid(static_cast<unsigned>(ix->nb), static_cast<unsigned>(m));
    tfidf_scores_kernel<<<grid, TF_THREADS, 0, st>>>(n, ix->nb, ix->blk_off, ix->post_doc, ix->post_val, d_ptr.as<int64_t>() + q0,
                                                    d_term.as<int32_t>(), d_val.as<double>(), dk.Current(), dv.Current());
    ARB_LAUNCH_OK();
    if (k <= SEL_MAX_K) {
      // np.argsort(similarities)[::-1] (retriever.py:980): descending similarity, the LARGER index first among equal
      // ones -- the keys are stored in reverse doc order, ties are taken in position order
      tfidf_select_kernel<<<static_cast<unsigned>(m), SEL_THREADS, 0, st>>>(dk.Current(), n, k, o_s.as<double>() + q0 * k,
                                                                          o_i.as<int64_t>() + q0 * k);
      ARB_LAUNCH_OK();
      continue;
    }
    // k beyond the select kernel's list: a stable segmented radix sort of all scores (library sort; same order)
    ARB_CUDA(cub::DeviceSegmentedRadixSort::SortPairsDescending(tmp.p, tmp_bytes, dk, dv, m * n, static_cast<int>(m),
                                                                seg.as<int64_t>(), seg.as<int64_t>() + 1, 0, 64, st));
    count_launch();
    tfidf_emit_kernel<<<static_cast<unsigned>((m * k + 255) / 256), 256, 0, st>>>(
        dk.Current(), dv.Current(), m, n, k, o_s.as<double>() + q0 * k, o_i.as<int64_t>() + q0 * k);
    ARB_LAUNCH_OK();
  }
  ARB_CUDA(cudaMemcpyAsync(out_scores, o_s.p, nq * k * 8, cudaMemcpyDeviceToHost, st));
  ARB_CUDA(cudaMemcpyAsync(out_ids, o_i.p, nq * k * 8, cudaMemcpyDeviceToHost, st));
  ARB_CUDA(cudaStreamSynchronize(st));
  return ARB_OK;
}
