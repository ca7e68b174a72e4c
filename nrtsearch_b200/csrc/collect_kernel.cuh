// Kernels either side of the top-k: the second pass of QueryRescorer (a query evaluated on a given hit list), the
// fetch phase on doc-value columns, and the aggregating "additional collectors" (terms / min / max / sum).
//   reference: src/main/java/com/yelp/nrtsearch/server/rescore/QueryRescore.java:39-57 (Lucene QueryRescorer.rescore),
//              .../handler/SearchHandler.java:397-522 (fetch: FillDocsTask / LoadedDocValues),
//              .../search/collectors/additional/{Int,Long,Float,Double}TermsCollectorManager.java, Max/Min/SumCollectorManager.java,
//              fan-out at .../search/SearchCollectorManager.java:192-198.
#pragma once
#include "bool_kernel.cuh"
#include "../../include/nrtgpu.h"

namespace nrtgpu {

// exact tf of (term clause, doc): dense byte plane when the term has one, else a binary search of its postings
__device__ __forceinline__ float term_freq_of(const DevIndexView& ix, const DevClause& c, int32_t doc, bool* present) {
  if (c.plane >= 0 && ix.dense_tf) {
    const uint32_t b = ix.dense_tf[(size_t)c.plane * (size_t)ix.dense_stride + doc];
    *present = b != 0;
    if (b != 255u) return (float)b;
    return exact_freq_slow<uint32_t>(ix, c, doc);
  }
  const int32_t* docs = ix.post_docs + c.post_base;
  int lo = 0, hi = c.n_post;
  while (lo < hi) { const int m = (lo + hi) >> 1; if (__ldg(docs + m) < doc) lo = m + 1; else hi = m; }
  *present = lo < c.n_post && __ldg(docs + lo) == doc;
  if (!*present) return 0.0f;
  const uint32_t b = ix.post_f8[c.post_base + lo];
  if (b != 255u) return (float)b;
  return exact_freq_slow<uint32_t>(ix, c, doc);
}

// One flat BooleanQuery on one doc: Lucene BooleanScorerSupplier semantics (conjunction / disjunction sums in double,
// ReqOptSumScorer float add when minShouldMatch == 0), identical to the top-k kernels' clause evaluation.
__device__ __forceinline__ bool eval_query_on_doc(const DevIndexView& ix, const DevQuery& q, const DevClause* __restrict__ cl,
                                                  int32_t doc, float* out_score) {
  if (q.empty) return false;
  if (ix.live_bits && !((ix.live_bits[doc >> 5] >> (doc & 31)) & 1u)) return false;
  double must_sum = 0.0, should_sum = 0.0;
  int n_should = 0;
  for (int i = 0; i < q.n_clauses; ++i) {
    const DevClause& c = cl[i];
    bool present;
    float s = 0.0f;
    if (c.kind == NRTGPU_TERM) {
      const float f = term_freq_of(ix, c, doc, &present);
      if (present && c.scoring) {
        const uint8_t* nrm = ix.norms[c.field];
        const uint32_t nb = nrm ? (uint32_t)nrm[doc] : 1u;
        s = bm25_score(c.weight, f, ix.caches[c.field * 256 + nb]);
      }
    } else if (c.kind == NRTGPU_RANGE_I64) {
      present = range_matches(ix, c.col, doc, c.lo, c.hi);
      s = c.weight;
    } else {
      present = true;
      s = c.weight;
    }
    if (!present) {
      if (c.occur == NRTGPU_MUST || c.occur == NRTGPU_FILTER) return false;
      continue;
    }
    switch (c.occur) {
      case NRTGPU_MUST: must_sum += (double)s; break;
      case NRTGPU_FILTER: break;
      case NRTGPU_SHOULD: should_sum += (double)s; ++n_should; break;
      default: return false;   // MUST_NOT present
    }
  }
  if (n_should < q.need_should) return false;
  float score;
  if (q.n_req == 0) score = (float)should_sum;
  else {
    const float req = (float)must_sum;
    if (n_should == 0) score = req;
    else {
      const float opt = (float)should_sum;
      score = (q.msm > 0) ? (float)((double)req + (double)opt) : __fadd_rn(req, opt);
    }
  }
  *out_score = score;
  return true;
}

// QueryRescorer second pass: query q on its own first-pass hits
struct ScoreDocsLaunch {
  DevIndexView ix;
  const DevClause* clauses; const DevQuery* queries;
  int32_t nq, n_hits;
  const int32_t* docs;     // [nq][n_hits] global doc ids
  const int32_t* counts;   // [nq] or NULL
  uint8_t* out_matches; float* out_scores;
};

__global__ void score_docs_kernel(ScoreDocsLaunch L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.nq * L.n_hits) return;
  const int q = i / L.n_hits, r = i % L.n_hits;
  uint8_t m = 0; float s = 0.0f;
  if (r < (L.counts ? L.counts[q] : L.n_hits)) {
    const int64_t local = (int64_t)L.docs[i] - L.ix.doc_base;
    if (local >= 0 && local < L.ix.n_docs) {
      const DevQuery dq = L.queries[q];
      float sc;
      if (eval_query_on_doc(L.ix, dq, L.clauses + dq.clause_begin, (int32_t)local, &sc)) { m = 1; s = sc; }
    }
  }
  L.out_matches[i] = m; L.out_scores[i] = s;
}

// fetch phase: doc values of n_cols columns for n docs
struct FetchLaunch {
  DevIndexView ix;
  const int32_t* col_ids; int32_t n_cols;
  const int32_t* docs; int32_t n;   // global doc ids
  int64_t* out_values;              // [n_cols][n]
  uint8_t* out_has;                 // [n_cols][n]
};
__global__ void fetch_columns_kernel(FetchLaunch F) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)F.n_cols * F.n) return;
  const int c = F.col_ids[i / F.n];
  const int64_t local = (int64_t)F.docs[i % F.n] - F.ix.doc_base;
  int64_t v = 0; uint8_t h = 0;
  if (local >= 0 && local < F.ix.n_docs) {
    const uint8_t* has = F.ix.col_has[c];
    h = (!has || has[local]) ? 1 : 0;
    if (h) v = F.ix.col32[c] ? (int64_t)F.ix.col32[c][local] : F.ix.col64[c][local];
  }
  F.out_values[i] = v; F.out_has[i] = h;
}

// ---- aggregations over ALL matching docs of every query (ScoreMode.COMPLETE, as RelevanceCollector.java:55-62 forces)
// value types of a column (how the sortable long maps to the double the Min/Max/Sum collectors see)
enum { kAggInt = 0, kAggFloat = 1, kAggDouble = 2 };
__device__ __forceinline__ double agg_value(int64_t v, int value_type) {
  if (value_type == kAggFloat) { int32_t b = (int32_t)v; b ^= (b >> 31) & 0x7fffffff; return (double)__int_as_float(b); }   // NumericUtils.sortableIntToFloat
  if (value_type == kAggDouble) { long long b = v; b ^= (b >> 63) & 0x7fffffffffffffffll; return __longlong_as_double(b); }
  return (double)v;
}

struct AggSpecDev {
  int32_t kind;        // NRTGPU_AGG_*
  int32_t column, value_type;
  int32_t n_buckets;   // terms: distinct values of the column
  unsigned int* counts;        // terms: [nq][n_buckets]
  unsigned long long* dvals;   // min / max: ordered-double bits [nq]; sum: double bits [nq] (atomicAdd(double))
};
constexpr int kMaxAggs = 8;
struct AggLaunch {
  AggSpecDev a[kMaxAggs];
  int32_t n_aggs;
  const uint32_t* codes[kMaxAggs];   // terms: sort codes of the column (bucket = code / 2 - 1)
};

__device__ __forceinline__ unsigned long long double_to_ordered(double d) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double ordered_to_double(unsigned long long u) {
  const unsigned long long b = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
  double d;
#ifdef __CUDA_ARCH__
  d = __longlong_as_double((long long)b);
#else
  memcpy(&d, &b, sizeof(d));
#endif
  return d;
}

// called by the posting kernels for every matching doc of query q
__device__ __forceinline__ void agg_collect(const AggLaunch& A, const DevIndexView& ix, int q, int32_t doc) {
  for (int i = 0; i < A.n_aggs; ++i) {
    const AggSpecDev& s = A.a[i];
    const uint8_t* has = ix.col_has[s.column];
    if (has && !has[doc]) continue;   // LoadedDocValues.size() == 0: nothing to collect for this doc
    if (s.kind == NRTGPU_AGG_TERMS) {
      const uint32_t code = A.codes[i][doc];
      if (code) atomicAdd(&s.counts[(size_t)q * s.n_buckets + (code >> 1) - 1], 1u);
    } else {
      const int64_t raw = ix.col32[s.column] ? (int64_t)ix.col32[s.column][doc] : ix.col64[s.column][doc];
      const double v = agg_value(raw, s.value_type);
      if (s.kind == NRTGPU_AGG_MAX) atomicMax(&s.dvals[q], double_to_ordered(v));
      else if (s.kind == NRTGPU_AGG_MIN) atomicMin(&s.dvals[q], double_to_ordered(v));
      else atomicAdd(reinterpret_cast<double*>(&s.dvals[q]), v);
    }
  }
}

// terms aggregation result of one query: the `size` buckets with the largest (or smallest) counts
// (TermsCollectorManager.fillBucketResultByCount :430-480), bucket keys as column values
struct AggTermsLaunch {
  const unsigned int* counts; int32_t n_buckets, nq, size, order_desc;
  const uint64_t* distinct;   // sorted distinct values (sortable u64) of the column
  int64_t* out_keys; int32_t* out_counts;   // [nq][size]
  int32_t* out_n;             // [nq] buckets returned
  int32_t* out_total_buckets; // [nq] non-empty buckets
  long long* out_other;       // [nq] docs counted in buckets not returned
};
constexpr int kAggChunk = 2048;
__global__ void __launch_bounds__(256) agg_terms_topk_kernel(AggTermsLaunch T) {
  __shared__ uint64_t keys[2 * kAggChunk];
  __shared__ unsigned long long sh_sum;
  __shared__ int sh_nonzero;
  const int q = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) { sh_sum = 0ull; sh_nonzero = 0; }
  __syncthreads();
  const unsigned int* row = T.counts + (size_t)q * T.n_buckets;
  int have = 0;
  unsigned long long my_sum = 0; int my_nz = 0;
  for (int base = 0; base < T.n_buckets; base += kAggChunk) {
    for (int i = tid; i < kAggChunk; i += 256) {
      const int bkt = base + i;
      uint64_t k = 0ull;
      if (bkt < T.n_buckets) {
        const unsigned int c = row[bkt];
        if (c) {
          ++my_nz; my_sum += c;
          const uint32_t hi = T.order_desc ? c : ~c;                  // larger key = earlier bucket
          k = ((uint64_t)hi << 32) | (uint32_t)(~(uint32_t)bkt);        // ties: smaller value first (the reference leaves ties unordered)
        }
      }
      keys[have + i] = k;
    }
    const int n = have + kAggChunk;
    const int m = next_pow2(n);
    for (int i = n + tid; i < m; i += 256) keys[i] = 0ull;
    __syncthreads();
    block_bitonic_sort_desc(keys, m);
    have = min(T.size, kAggChunk);
    __syncthreads();
  }
  atomicAdd(&sh_sum, my_sum); atomicAdd(&sh_nonzero, my_nz);
  __syncthreads();
  int n_out = 0;
  unsigned long long shown = 0;
  for (int i = 0; i < have; ++i) if (keys[i]) ++n_out; else break;   // (uniform: every thread scans the same smem)
  for (int i = tid; i < T.size; i += 256) {
    int64_t key = 0; int32_t cnt = 0;
    if (i < n_out) {
      const uint32_t hi = (uint32_t)(keys[i] >> 32), bkt = ~(uint32_t)keys[i];
      cnt = (int32_t)(T.order_desc ? hi : ~hi);
      key = (int64_t)(T.distinct[bkt] ^ 0x8000000000000000ull);
    }
    T.out_keys[(size_t)q * T.size + i] = key; T.out_counts[(size_t)q * T.size + i] = cnt;
  }
  if (tid == 0) {
    for (int i = 0; i < n_out; ++i) { const uint32_t hi = (uint32_t)(keys[i] >> 32); shown += T.order_desc ? hi : ~hi; }
    T.out_n[q] = n_out; T.out_total_buckets[q] = sh_nonzero; T.out_other[q] = (long long)(sh_sum - shown);
  }
}

}  // namespace nrtgpu
