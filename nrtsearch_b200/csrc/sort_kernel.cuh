// Sort-by-field top-k (TopFieldCollector semantics; reference
// src/main/java/com/yelp/nrtsearch/server/search/collectors/SortFieldCollector.java:44-105, numeric sort fields
// .../field/NumberFieldDef.java:266-278, missing values IntFieldDef.java:103 / LongFieldDef.java:103 / ...).
//
// The top-k machinery of the posting kernels orders 64-bit keys (hi word desc, then ~doc desc = doc asc). A sorted
// search swaps the key function: hi = an ORDER-PRESERVING 32-bit code of the doc's sort value. The codes are index-time
// data, one uint32 per doc and column: the column's distinct values are sorted once on the device and value number i
// (0-based) gets code 2i + 2; code 2i + 1 stands for "between value i-1 and value i" (the code of a value the column
// does not hold: a missing_value such as Long.MIN_VALUE, or the searchAfter value of another shard). So any int64 can
// be compared with every doc exactly, ties included (a doc WITHOUT a value takes the code of missing_value at query
// time, and therefore ties with docs that really hold that value, as Lucene's comparator does).
#pragma once
#include <thrust/copy.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include "bool_kernel.cuh"

namespace nrtgpu {

__host__ __device__ __forceinline__ uint64_t sortable_u64(int64_t v) { return (uint64_t)v ^ 0x8000000000000000ull; }

struct SortNonZero { __host__ __device__ bool operator()(uint8_t x) const { return x != 0; } };

// keys[i] = sortable value of doc idx[i] (idx = the docs that have a value)
__global__ void sort_col_keys_kernel(const int64_t* __restrict__ c64, const int32_t* __restrict__ c32, const int32_t* __restrict__ idx,
                                     int32_t n, uint64_t* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t d = idx[i];
  keys[i] = sortable_u64(c32 ? (int64_t)c32[d] : c64[d]);
}
__global__ void sort_mark_kernel(const uint64_t* __restrict__ keys, int32_t n, int32_t* __restrict__ mark) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  mark[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
// rank[i] = 1-based number of the distinct value of sorted position i
__global__ void sort_scatter_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ idx, const int32_t* __restrict__ rank,
                                    int32_t n, uint32_t* __restrict__ codes, uint64_t* __restrict__ distinct) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t r = rank[i] - 1;
  codes[idx[i]] = 2u * (uint32_t)r + 2u;
  if (i == 0 || keys[i] != keys[i - 1]) distinct[r] = keys[i];
}

// order-preserving code of an arbitrary value against a column's sorted distinct values
__device__ __forceinline__ uint32_t sort_code_of(const uint64_t* __restrict__ distinct, int32_t n_distinct, int64_t v) {
  const uint64_t k = sortable_u64(v);
  int lo = 0, hi = n_distinct;
  while (lo < hi) { const int m = (lo + hi) >> 1; if (distinct[m] < k) lo = m + 1; else hi = m; }
  return (lo < n_distinct && distinct[lo] == k) ? 2u * (uint32_t)lo + 2u : 2u * (uint32_t)lo + 1u;
}

struct SortAfterLaunch {
  DevQuery* queries; int32_t nq;
  const int32_t* after_docs;      // [nq] global doc ids (nrtgpu_query.after_doc)
  const int64_t* after_values;    // [nq]
  int32_t kind, reverse, doc_base, n_docs;
  const uint64_t* distinct; int32_t n_distinct;
  int64_t missing_value;
  uint32_t* missing_code;         // [1] out: code of missing_value
};

__device__ __forceinline__ uint32_t sort_hi(int kind, int reverse, uint32_t code, int32_t doc) {
  if (kind == NRTGPU_SORT_DOCID) return reverse ? (uint32_t)doc + 1u : 0x7fffffffu;
  return reverse ? code : ~code;
}

// patches DevQuery::after_key of every query with searchAfter: a hit qualifies iff key < after_key
__global__ void sort_after_kernel(SortAfterLaunch S) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q == 0 && S.kind == NRTGPU_SORT_COLUMN) *S.missing_code = sort_code_of(S.distinct, S.n_distinct, S.missing_value);
  if (q >= S.nq || !S.queries[q].has_after) return;
  const int64_t local = (int64_t)S.after_docs[q] - S.doc_base;
  uint64_t key;
  if (S.kind == NRTGPU_SORT_DOCID && S.reverse) {
    key = local < 0 ? 0ull : (local >= S.n_docs ? 0xffffffffffffffffull : ((uint64_t)((uint32_t)local + 1u) << 32));
  } else {
    uint32_t code = 0; bool exact = true;
    if (S.kind == NRTGPU_SORT_COLUMN) { code = sort_code_of(S.distinct, S.n_distinct, S.after_values[q]); exact = (code & 1u) == 0u; }
    const uint32_t hi = sort_hi(S.kind, S.reverse, code, 0);
    if (!exact) key = (uint64_t)hi << 32;                                   // no doc holds the value: the doc part is moot
    else if (local < 0) key = ((uint64_t)hi + 1ull) << 32;                  // every tied doc here follows afterDoc
    else if (local >= S.n_docs) key = (uint64_t)hi << 32;                   // every tied doc here precedes it
    else key = ((uint64_t)hi << 32) | (uint32_t)(~(uint32_t)local);
  }
  S.queries[q].after_key = key;
}

// FieldDoc.fields[0] of the final hits
struct SortValuesLaunch {
  const int32_t* docs; const int32_t* counts; int32_t nq, top_k, doc_base, kind;
  const int64_t* c64; const int32_t* c32; const uint8_t* has; int64_t missing_value;
  int64_t* out_values; float* out_scores;
};
__global__ void sort_values_kernel(SortValuesLaunch S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.nq * S.top_k) return;
  const int q = i / S.top_k, r = i % S.top_k;
  if (S.out_scores) S.out_scores[i] = __int_as_float(0x7fc00000);   // NaN: TopFieldCollector does not track scores
  if (r >= S.counts[q]) { S.out_values[i] = 0; return; }
  const int32_t g = S.docs[i], d = g - S.doc_base;
  if (S.kind == NRTGPU_SORT_DOCID) { S.out_values[i] = g; return; }
  if (S.has && !S.has[d]) { S.out_values[i] = S.missing_value; return; }
  S.out_values[i] = S.c32 ? (int64_t)S.c32[d] : S.c64[d];
}

// index time: codes + sorted distinct values of one column (device pointers; codes zeroed by the caller: 0 = no value)
inline int sort_codes_build(const int64_t* c64, const int32_t* c32, const uint8_t* has, int32_t n, uint32_t* codes,
                            uint64_t* keys_tmp, int32_t* idx_tmp, int32_t* rank_tmp, uint64_t* distinct, int32_t* n_distinct_out) {
  if (n <= 0) { *n_distinct_out = 0; return NRTGPU_OK; }
  // the docs that have a value, then their sortable keys, sorted
  int32_t n_has = n;
  if (has) n_has = (int32_t)(thrust::copy_if(thrust::device, thrust::counting_iterator<int32_t>(0), thrust::counting_iterator<int32_t>(n), has, idx_tmp, SortNonZero()) - idx_tmp);
  else thrust::sequence(thrust::device, idx_tmp, idx_tmp + n);
  if (n_has <= 0) { *n_distinct_out = 0; return NRTGPU_OK; }
  sort_col_keys_kernel<<<(unsigned)((n_has + 255) / 256), 256>>>(c64, c32, idx_tmp, n_has, keys_tmp);
  NRT_CUDA_TRY(cudaGetLastError());
  thrust::sort_by_key(thrust::device, keys_tmp, keys_tmp + n_has, idx_tmp);
  sort_mark_kernel<<<(unsigned)((n_has + 255) / 256), 256>>>(keys_tmp, n_has, rank_tmp);
  NRT_CUDA_TRY(cudaGetLastError());
  thrust::inclusive_scan(thrust::device, rank_tmp, rank_tmp + n_has, rank_tmp);
  sort_scatter_kernel<<<(unsigned)((n_has + 255) / 256), 256>>>(keys_tmp, idx_tmp, rank_tmp, n_has, codes, distinct);
  NRT_CUDA_TRY(cudaGetLastError());
  int32_t last = 0;
  NRT_CUDA_TRY(cudaMemcpy(&last, rank_tmp + n_has - 1, sizeof(int32_t), cudaMemcpyDeviceToHost));
  *n_distinct_out = last;
  return NRTGPU_OK;
}

}  // namespace nrtgpu
