// Shared device/host helpers for the nrtgpu kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace nrtgpu {

// ---- error plumbing (thread-local message, returned through nrtgpu_last_error) ----
void set_error(const std::string& msg);
#define NRT_CUDA_TRY(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::nrtgpu::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));               \
      return (_e == cudaErrorMemoryAllocation) ? NRTGPU_ERR_OOM : NRTGPU_ERR_CUDA;           \
    }                                                                                        \
  } while (0)

// ---- total order on hits: (score desc, doc asc)  <=>  key desc ----
// reference: src/main/java/org/apache/lucene/search/LazyQueueTopScoreDocCollector.java:129-143
// key = ordered(score) << 32 | ~doc ; all keys of real hits are > 0, so 0 is the "empty" sentinel.
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, int32_t doc) {
  return ((uint64_t)float_to_ordered(score) << 32) | (uint32_t)(~(uint32_t)doc);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return ordered_to_float((uint32_t)(k >> 32)); }
__host__ __device__ __forceinline__ int32_t key_doc(uint64_t k) { return (int32_t)(~(uint32_t)k); }

#ifdef __CUDACC__
// BM25 term score exactly as Lucene's BM25Scorer.score (float ops, round-to-nearest, never fused):
//   weight - weight / (1f + freq * cache[norm])
__device__ __forceinline__ float bm25_score(float weight, float freq, float norm_inverse) {
  float x = __fmul_rn(freq, norm_inverse);
  x = __fadd_rn(1.0f, x);
  x = __fdiv_rn(weight, x);
  return __fsub_rn(weight, x);
}

// In-place bitonic sort (descending) of n (power of two) 64-bit keys in shared memory by the whole CTA.
__device__ __forceinline__ void block_bitonic_sort_desc(uint64_t* a, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        int hi = lo | j;
        bool desc = ((lo & k) == 0);
        uint64_t x = a[lo], y = a[hi];
        if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}
#endif

}  // namespace nrtgpu
