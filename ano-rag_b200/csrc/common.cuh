// Shared helpers for libanorag_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include "../../include/anorag_b200.h"

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL)
#error "libanorag_b200 is written for sm_100a only (compile with -gencode arch=compute_100a,code=sm_100a)"
#endif

namespace arb {

void set_error(const std::string& msg);
int  cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(int n = 1);
void configure_pool_once();

#define ARB_CUDA(expr)                                                         \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ::arb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define ARB_CHECK(cond, code, msg)                                             \
  do {                                                                         \
    if (!(cond)) { ::arb::set_error(msg); return (code); }                     \
  } while (0)

#define ARB_LAUNCH_OK()                                                        \
  do {                                                                         \
    ::arb::count_launch();                                                     \
    ARB_CUDA(cudaGetLastError());                                              \
  } while (0)

// Stream-ordered scratch: freed on the same stream when the guard dies, so
// concurrent searches on one handle never share scratch.
struct AsyncBuf {
  void* p = nullptr;
  cudaStream_t s = nullptr;
  cudaError_t alloc(size_t bytes, cudaStream_t stream) {
    configure_pool_once();
    s = stream;
    if (bytes == 0) bytes = 16;
    return cudaMallocAsync(&p, bytes, stream);
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
  ~AsyncBuf() { if (p) cudaFreeAsync(p, s); }
  AsyncBuf() = default;
  AsyncBuf(const AsyncBuf&) = delete;
  AsyncBuf& operator=(const AsyncBuf&) = delete;
};

int num_sms();

// Optional per-thread kernel timing (bench.py roofline): when enabled, the dominant
// kernels are bracketed by CUDA events on their launch stream.
enum { PROF_DENSE_TOPK = 0, PROF_BM25_TOPK = 1, PROF_FUSE = 2, PROF_DENSE_MERGE = 3, PROF_BM25_SPARSE = 4,
       PROF_BM25_PLANES = 5, PROF_BM25_FINAL = 6, PROF_IVF_COARSE = 7, PROF_KINDS = 8 };
struct ProfileScope {
  int kind; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr; bool on;
  ProfileScope(int kind, cudaStream_t st);
  ~ProfileScope();
};

// order-preserving map fp32 -> uint32 (larger float <=> larger key); 0 is never
// produced by a finite value, so 0 doubles as "empty".
__host__ __device__ __forceinline__ uint32_t f32_key(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  uint32_t b; memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float key_f32(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f; memcpy(&f, &b, 4); return f;
#endif
}
__device__ __forceinline__ unsigned long long f64_key(double d) {
  unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

}  // namespace arb
