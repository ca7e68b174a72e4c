// Batched BooleanQuery execution over HBM-resident postings: window engine + slice merge.
//
// Replaces, for a whole batch of queries at once, the per-query chain
//   IndexSearcher.search -> BooleanWeight.bulkScorer -> (MaxScoreBulkScorer | ConjunctionDISI) ->
//   TermScorer/BM25Scorer -> TopScoreDocCollector
// that the reference drives from src/main/java/com/yelp/nrtsearch/server/handler/SearchHandler.java:1412.
//
// Work item = (doc slice, query). A CTA sweeps its slice in windows of W docs:
//   pass 1 (scatter): every posting of every term clause in the window stores its saturated tf byte
//                     into byte `slot` of a W-entry shared-memory word array (no atomics: doc ids
//                     are unique inside one posting list and the passes are barrier separated);
//   pass 2 (emit):    the postings of the DRIVER clauses are re-read; the lowest driver clause present
//                     in a doc's word "owns" the doc, clears the word, evaluates the boolean
//                     constraints, scores the matched clauses with Lucene's BM25 float formula, sums
//                     in double in clause order, and offers (score, doc) to the CTA's candidate
//                     buffer if it beats the query's running threshold theta;
//   pass 3 (clean):   non-driver clauses clear the words pass 2 did not visit.
// Smem cost is proportional to postings, not to W. theta is a 64-bit (score, ~doc) key shared by all
// slices of a query through one global atomicMax word -- the device analogue of the reference's
// LazyMaxScoreAccumulator (src/main/java/org/apache/lucene/search/LazyMaxScoreAccumulator.java:21-70),
// used here only to drop hits that provably cannot enter the top-k (results stay exact).
#pragma once
#include "common.cuh"

namespace nrtgpu {

constexpr int kMaxClauses = 16;     // clauses per flat BooleanQuery on the GPU path
constexpr int kMaxTermSlots = 8;    // term clauses per query (u32 words: 4, u64 words: 8)
constexpr int kWindowDocs = 16384;  // W
constexpr int kSliceWindows = 64;   // windows per work item  => 1,048,576 docs per slice
constexpr int kCandCap = 4096;      // candidate buffer (keys) per CTA, power of two
constexpr int kThreads = 512;
constexpr int kMaxTopK = 1024;

struct DevClause {
  int64_t post_base;  // offset of the term's postings in post_docs / post_f8
  int32_t n_post;
  int32_t occur;
  int32_t kind;
  int32_t slot;       // byte index inside the window word (term clauses), -1 otherwise
  int32_t field;      // text field (norms + cache) for term clauses
  int32_t col;        // doc-value column for range clauses
  float weight;       // boost*idf (term) or constant score = boost (range / match-all)
  int32_t scoring;    // 1 if the clause contributes to the score (MUST / SHOULD)
  float ub;           // term clauses: largest score of any posting of the list (index-time max of tf*cache[norm])
  int32_t plane;      // term clauses: dense tf plane of the term (DevIndexView::dense_tf), -1 if the term has none
  int32_t gran_row;   // term clauses: row of the index-time granule offset table (DevIndexView::gran_tab), -1 if none
  int32_t pad_;
  int64_t lo, hi;
};

struct DevQuery {
  int32_t clause_begin, n_clauses;
  int32_t n_term;          // number of term clauses (= slots used)
  int32_t n_req;           // MUST + FILTER clauses (all kinds)
  int32_t need_should;     // minimum matching SHOULD clauses
  int32_t msm;             // minimumNumberShouldMatch as given
  uint32_t req_term_mask;  // bit s set: term slot s is MUST/FILTER
  uint32_t not_term_mask;  // bit s set: term slot s is MUST_NOT
  uint32_t driver_mask;    // bit s set: term slot s drives pass 2
  int32_t dense_driver;    // 1: iterate every doc of the window instead of driver postings
  int32_t has_non_driver;  // 1: some term slot is not a driver (pass 3 needed)
  int32_t has_nonterm;     // 1: range / match-all clauses present
  int32_t empty;           // 1: can match nothing
  int32_t has_after;
  uint32_t must_term_mask;    // bit s set: term slot s is MUST (scores into the required sum)
  uint32_t should_term_mask;  // bit s set: term slot s is SHOULD
  int32_t nonterm_scoring;    // 1: a range / match-all clause is MUST or SHOULD (contributes a constant score)
  int32_t single_field;       // >= 0: every term clause reads this text field's norms; -1: mixed
  uint64_t after_key;
};

struct DevIndexView {
  int32_t n_docs;
  int32_t doc_base;
  const int32_t* post_docs;
  const uint8_t* post_f8;        // min(freq, 255)
  const int64_t* exc_pos;        // sorted global posting indices with freq >= 255
  const int32_t* exc_freq;
  int32_t n_exc;
  const uint8_t* const* norms;   // [n_fields] device pointers (NULL = omitNorms)
  const float* caches;           // [n_fields][256]
  const int64_t* const* col64;   // [n_columns] (NULL if stored as int32)
  const int32_t* const* col32;   // [n_columns] (NULL if stored as int64)
  const uint8_t* const* col_has; // [n_columns] (NULL = all)
  const int64_t* const* colmv_off;  // [n_columns] multi-valued columns (SORTED_NUMERIC): doc d holds colmv_val[c][off[d] .. off[d + 1]),
  const int64_t* const* colmv_val;  //              ascending; NULL entry = single-valued column
  const uint32_t* live_bits;     // bitmap or NULL
  const uint32_t* gran_tab;      // [n_rows][n_gran + 1] postings of the term below each stream-kernel granule (skip data)
  int32_t n_gran;
  const uint8_t* dense_tf;       // [n_planes][dense_stride] min(freq, 255) per doc for the densest terms (0 = absent)
  int64_t dense_stride;
  const uint8_t* dense_tf2;      // [n_planes][dense_stride / 4] min(freq, 3) in 2 bits per doc: the planes the probe kernel gathers
                                 // (a quarter of the L2 / DRAM footprint of the byte planes; 3 = "three or more")
};

struct BoolLaunch {
  DevIndexView ix;
  const DevClause* clauses;
  const DevQuery* queries;
  const int32_t* work_query;   // [n_work] query index per work item (slice-major order)
  const int32_t* work_slice;   // [n_work]
  int32_t n_work;
  int32_t n_slices;
  int32_t top_k;
  uint64_t* theta;             // [nq] running k-th best key (0 = none yet)
  unsigned long long* total_hits;  // [nq]
  uint64_t* slice_keys;        // [nq][n_slices][top_k]
  int32_t* slice_cnt;          // [nq][n_slices]
};

// numeric range clause on one doc (IndexOrDocValuesQuery's doc-values side, reference IntFieldDef.java:124-158 inclusive
// bounds): single-valued column = the value is in [lo, hi]; multi-valued (SortedNumericDocValuesRangeQuery) = ANY value is
__device__ __forceinline__ bool range_matches(const DevIndexView& ix, int col, int32_t doc, int64_t lo, int64_t hi) {
  const int64_t* off = ix.colmv_off ? ix.colmv_off[col] : nullptr;
  if (off) {
    const int64_t* v = ix.colmv_val[col];
    int64_t a = off[doc], b = off[doc + 1];
    while (a < b) { const int64_t m = (a + b) >> 1; if (v[m] < lo) a = m + 1; else b = m; }   // values of a doc are sorted
    return a < off[doc + 1] && v[a] <= hi;
  }
  const uint8_t* has = ix.col_has[col];
  if (has && !has[doc]) return false;
  const int64_t x = ix.col32[col] ? (int64_t)__ldg(ix.col32[col] + doc) : __ldg(ix.col64[col] + doc);
  return x >= lo && x <= hi;
}

template <typename SlotT>
struct SlotTraits;
template <> struct SlotTraits<uint32_t> { static constexpr int kSlots = 4; };
template <> struct SlotTraits<uint64_t> { static constexpr int kSlots = 8; };

template <typename SlotT>
struct BoolSmem {
  SlotT slots[kWindowDocs];
  uint64_t cand[kCandCap];
  uint32_t bounds[kMaxTermSlots][kSliceWindows + 1];
  float cache[kMaxTermSlots][256];
  DevClause cl[kMaxClauses];
  DevQuery q;
  int cand_count;
  unsigned long long theta;
};

template <typename SlotT>
__device__ __forceinline__ uint32_t presence_mask(SlotT s) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < SlotTraits<SlotT>::kSlots; ++i) m |= (((s >> (8 * i)) & 0xff) != 0 ? 1u : 0u) << i;
  return m;
}

// exact tf of posting (clause c, doc) when the byte saturated: find the posting, then the exception list
template <typename SlotT>
__device__ __noinline__ float exact_freq_slow(const DevIndexView& ix, const DevClause& c, int32_t doc) {
  const int32_t* docs = ix.post_docs + c.post_base;
  int lo = 0, hi = c.n_post;
  while (lo < hi) { int m = (lo + hi) >> 1; if (docs[m] < doc) lo = m + 1; else hi = m; }
  int64_t gp = c.post_base + lo;
  int a = 0, b = ix.n_exc;
  while (a < b) { int m = (a + b) >> 1; if (ix.exc_pos[m] < gp) a = m + 1; else b = m; }
  if (a < ix.n_exc && ix.exc_pos[a] == gp) return (float)ix.exc_freq[a];
  return 255.0f;
}

// Evaluate the boolean constraints + score for one candidate doc. Returns false if the doc does not match.
// Score combination follows Lucene's BooleanScorerSupplier: conjunction / disjunction sums are double,
// required+optional is ReqOptSumScorer's float add (msm == 0) or ConjunctionScorer's double add (msm > 0).
template <typename SlotT>
__device__ __forceinline__ bool evaluate_doc(const DevIndexView& ix, const BoolSmem<SlotT>& sm, int32_t doc,
                                             SlotT slot, float* out_score) {
  const DevQuery& q = sm.q;
  uint32_t m = presence_mask<SlotT>(slot);
  if ((m & q.req_term_mask) != q.req_term_mask) return false;
  if (m & q.not_term_mask) return false;
  if (ix.live_bits && !((ix.live_bits[doc >> 5] >> (doc & 31)) & 1u)) return false;
  double must_sum = 0.0, should_sum = 0.0;
  int n_req = 0, n_should = 0;
  for (int i = 0; i < q.n_clauses; ++i) {
    const DevClause& c = sm.cl[i];
    bool present;
    float s = 0.0f;
    if (c.kind == NRTGPU_TERM) {
      uint32_t b = (uint32_t)((slot >> (8 * c.slot)) & 0xff);
      present = b != 0;
      if (present && c.scoring) {
        float f = (b == 255u) ? exact_freq_slow<SlotT>(ix, c, doc) : (float)b;
        const uint8_t* nrm = ix.norms[c.field];
        uint32_t nb = nrm ? (uint32_t)nrm[doc] : 1u;
        s = bm25_score(c.weight, f, sm.cache[c.slot][nb]);
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
      case NRTGPU_MUST: must_sum += (double)s; ++n_req; break;
      case NRTGPU_FILTER: ++n_req; break;
      case NRTGPU_SHOULD: should_sum += (double)s; ++n_should; break;
      default: return false;  // MUST_NOT present
    }
  }
  if (n_should < q.need_should) return false;
  float score;
  if (q.n_req == 0) score = (float)should_sum;
  else {
    float req = (float)must_sum;
    if (n_should == 0) score = req;
    else {
      float opt = (float)should_sum;
      score = (q.msm > 0) ? (float)((double)req + (double)opt) : __fadd_rn(req, opt);
    }
  }
  *out_score = score;
  return true;
}

// sort the candidate buffer, keep the best top_k, raise theta (local + global)
template <typename SlotT>
__device__ __forceinline__ void compact_candidates(BoolSmem<SlotT>& sm, int top_k, uint64_t* g_theta) {
  __syncthreads();
  int n = sm.cand_count;
  if (n > kCandCap) n = kCandCap;  // cannot happen (capacity invariant); defensive
  int m = next_pow2(n < 2 ? 2 : n);
  for (int i = n + threadIdx.x; i < m; i += blockDim.x) sm.cand[i] = 0ull;
  __syncthreads();
  block_bitonic_sort_desc(sm.cand, m);
  if (threadIdx.x == 0) {
    int keep = n < top_k ? n : top_k;
    sm.cand_count = keep;
    if (keep == top_k) {
      unsigned long long kth = sm.cand[top_k - 1];
      unsigned long long old = atomicMax((unsigned long long*)g_theta, kth);
      unsigned long long t = old > kth ? old : kth;
      if (t > sm.theta) sm.theta = t;
    } else {
      unsigned long long g = *(volatile unsigned long long*)g_theta;
      if (g > sm.theta) sm.theta = g;
    }
  }
  __syncthreads();
}

template <typename SlotT>
__global__ void __launch_bounds__(kThreads, 2) bool_window_kernel(BoolLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BoolSmem<SlotT>& sm = *reinterpret_cast<BoolSmem<SlotT>*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31;

  for (int wi = blockIdx.x; wi < L.n_work; wi += gridDim.x) {
    const int qi = L.work_query[wi];
    const int slice = L.work_slice[wi];
    __syncthreads();  // previous work item fully retired
    if (tid == 0) {
      sm.q = L.queries[qi];
      sm.cand_count = 0;
      sm.theta = *(volatile unsigned long long*)&L.theta[qi];
    }
    __syncthreads();
    const int ncl = sm.q.n_clauses;
    if (tid < ncl) sm.cl[tid] = L.clauses[sm.q.clause_begin + tid];
    for (int i = tid; i < kWindowDocs; i += kThreads) sm.slots[i] = 0;
    __syncthreads();
    // per-slot BM25 caches
    for (int i = tid; i < ncl * 256; i += kThreads) {
      int c = i >> 8;
      if (sm.cl[c].kind == NRTGPU_TERM) sm.cache[sm.cl[c].slot][i & 255] = L.ix.caches[sm.cl[c].field * 256 + (i & 255)];
    }
    const int32_t slice_base = slice * (kSliceWindows * kWindowDocs);
    int32_t slice_end = slice_base + kSliceWindows * kWindowDocs;
    if (slice_end > L.ix.n_docs || slice_end < 0) slice_end = L.ix.n_docs;
    const int nwin = (slice_end - slice_base + kWindowDocs - 1) / kWindowDocs;
    // posting bounds of every term clause at every window boundary (lower_bound over the full list)
    for (int i = tid; i < ncl * (kSliceWindows + 1); i += kThreads) {
      int c = i / (kSliceWindows + 1), w = i % (kSliceWindows + 1);
      if (sm.cl[c].kind != NRTGPU_TERM) continue;
      int64_t target64 = (int64_t)slice_base + (int64_t)w * kWindowDocs;
      int32_t target = target64 > (int64_t)slice_end ? slice_end : (int32_t)target64;
      const int32_t* docs = L.ix.post_docs + sm.cl[c].post_base;
      int lo = 0, hi = sm.cl[c].n_post;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(docs + mid) < target) lo = mid + 1; else hi = mid; }
      sm.bounds[sm.cl[c].slot][w] = (uint32_t)lo;
    }
    __syncthreads();

    unsigned long long my_hits = 0;
    int cand_ub = 0;  // CTA-uniform upper bound of cand_count
    const bool dense = sm.q.dense_driver != 0;
    const bool has_after = sm.q.has_after != 0;
    const uint64_t after_key = sm.q.after_key;

    for (int w = 0; w < nwin; ++w) {
      const int32_t wbase = slice_base + w * kWindowDocs;
      const int32_t wlen = min(kWindowDocs, slice_end - wbase);
      // ---------------- pass 1: scatter tf bytes
      bool any = dense;
      for (int c = 0; c < ncl; ++c) {
        if (sm.cl[c].kind != NRTGPU_TERM) continue;
        const int s = sm.cl[c].slot;
        const uint32_t b0 = sm.bounds[s][w], b1 = sm.bounds[s][w + 1];
        if (b1 > b0) any = true;
        const int32_t* docs = L.ix.post_docs + sm.cl[c].post_base;
        const uint8_t* f8 = L.ix.post_f8 + sm.cl[c].post_base;
        const bool scoring = sm.cl[c].scoring != 0;
        unsigned char* slot_bytes = reinterpret_cast<unsigned char*>(sm.slots);
        for (uint32_t p = b0 + tid; p < b1; p += kThreads) {
          int32_t d = docs[p] - wbase;
          unsigned char f = scoring ? f8[p] : (unsigned char)1;
          slot_bytes[(size_t)d * sizeof(SlotT) + s] = f;
        }
      }
      if (!any) continue;  // CTA-uniform: no postings in this window
      __syncthreads();
      // ---------------- pass 2: emit
      auto offer = [&](bool matched, int32_t doc, float score) {
        // converged call (all lanes of the warp)
        bool is_cand = false;
        uint64_t key = 0;
        if (matched) {
          ++my_hits;
          key = make_key(score, doc);
          is_cand = key > sm.theta && (!has_after || key < after_key);
        }
        unsigned bal = __ballot_sync(0xffffffffu, is_cand);
        if (bal) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&sm.cand_count, __popc(bal));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (is_cand) sm.cand[base + __popc(bal & ((1u << lane) - 1))] = key;
        }
      };
      auto round_end = [&]() {
        cand_ub += kThreads;
        if (cand_ub > kCandCap - kThreads) {
          __syncthreads();
          int n = sm.cand_count;
          if (n > kCandCap - kThreads) { compact_candidates(sm, L.top_k, &L.theta[qi]); n = sm.cand_count; }
          cand_ub = n;
        }
      };
      if (!dense) {
        for (int c = 0; c < ncl; ++c) {
          if (sm.cl[c].kind != NRTGPU_TERM) continue;
          const int s = sm.cl[c].slot;
          if (!((sm.q.driver_mask >> s) & 1u)) continue;
          // bytes of lower driver slots
          SlotT below = 0;
          for (int j = 0; j < s; ++j) if ((sm.q.driver_mask >> j) & 1u) below |= (SlotT)0xff << (8 * j);
          const SlotT own = (SlotT)0xff << (8 * s);
          const uint32_t b0 = sm.bounds[s][w], b1 = sm.bounds[s][w + 1];
          const int32_t* docs = L.ix.post_docs + sm.cl[c].post_base;
          for (uint32_t p0 = b0; p0 < b1; p0 += kThreads) {
            uint32_t p = p0 + tid;
            bool matched = false; int32_t doc = 0; float score = 0.0f;
            if (p < b1) {
              doc = docs[p];
              SlotT v = sm.slots[doc - wbase];
              if ((v & below) == 0 && (v & own) != 0) {
                sm.slots[doc - wbase] = 0;
                matched = evaluate_doc<SlotT>(L.ix, sm, doc, v, &score);
              }
            }
            offer(matched, doc, score);
            round_end();
          }
          __syncthreads();   // the words this list cleared are seen cleared by the next driver list (either order gave the same
                             // result -- the doc is skipped -- but the ordering is now explicit: racecheck-clean)
        }
      } else {
        for (int i0 = 0; i0 < wlen; i0 += kThreads) {
          int i = i0 + tid;
          bool matched = false; int32_t doc = wbase + i; float score = 0.0f;
          if (i < wlen) {
            SlotT v = sm.slots[i];
            if (v) sm.slots[i] = 0;
            matched = evaluate_doc<SlotT>(L.ix, sm, doc, v, &score);
          }
          offer(matched, doc, score);
          round_end();
        }
      }
      __syncthreads();
      // ---------------- pass 3: clear words of non-driver clauses
      if (!dense && sm.q.has_non_driver) {
        for (int c = 0; c < ncl; ++c) {
          if (sm.cl[c].kind != NRTGPU_TERM) continue;
          const int s = sm.cl[c].slot;
          if ((sm.q.driver_mask >> s) & 1u) continue;
          const uint32_t b0 = sm.bounds[s][w], b1 = sm.bounds[s][w + 1];
          const int32_t* docs = L.ix.post_docs + sm.cl[c].post_base;
          for (uint32_t p = b0 + tid; p < b1; p += kThreads) sm.slots[docs[p] - wbase] = 0;
        }
        __syncthreads();
      }
    }
    // ---------------- finish the work item
    compact_candidates(sm, L.top_k, &L.theta[qi]);
    const int keep = sm.cand_count;
    uint64_t* out = L.slice_keys + ((size_t)qi * L.n_slices + slice) * L.top_k;
    for (int i = tid; i < keep; i += kThreads) out[i] = sm.cand[i];
    if (tid == 0) L.slice_cnt[(size_t)qi * L.n_slices + slice] = keep;
    // total hits
    for (int o = 16; o > 0; o >>= 1) my_hits += __shfl_xor_sync(0xffffffffu, my_hits, o);
    if (lane == 0 && my_hits) atomicAdd(&L.total_hits[qi], my_hits);
  }
}

// ---- per-query merge of the slice lists (TopDocs.merge semantics: key order is total) ----
struct MergeLaunch {
  const uint64_t* slice_keys;  // [nq][n_lists][top_k]
  const int32_t* slice_cnt;    // [nq][n_lists]
  int32_t n_lists, top_k, nq;
  int32_t doc_base;
  int32_t* out_docs; float* out_scores; int32_t* out_counts;
  // optional: per-query totalHits / relation of the shard written next to the hits (the packed record of the one
  // all-gather of a multi-GPU step, nrtgpu_batch_bind_packed)
  const unsigned long long* total_hits; const int32_t* pruned; int32_t* terminated;
  long long terminate_after;                  // > 0: a query with more hits than this terminated early (TerminateAfterWrapper.java:150-158)
  long long* out_total; int32_t* out_flags;   // flags: bit 0 relation GREATER_THAN_OR_EQUAL_TO, bit 1 terminated early
  const unsigned long long* known_hits = nullptr;   // optional [nq]: docs known to match (reported totalHits = max(counted, known))
  const uint64_t* theta = nullptr;            // optional [nq]: the k-th best key some work item published (>= top_k keys are >= it):
                                              // smaller keys cannot be in the merged page and are dropped before the sort
};

constexpr int kMergeThreads = 256;
constexpr int kMergeCap = 4096;

__global__ void __launch_bounds__(kMergeThreads) merge_slices_kernel(MergeLaunch M) {
  __shared__ uint64_t keys[kMergeCap];
  __shared__ int32_t s_cnt[kMergeThreads];
  __shared__ int32_t s_nz[kMergeThreads];   // the non-empty lists of a chunk of list counts
  __shared__ int32_t s_off[kMergeThreads];
  __shared__ int32_t s_nnz;
  __shared__ int32_t s_fill;
  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  const uint64_t floor_key = M.theta ? M.theta[q] : 0ull;
  int have = 0;        // keys[0..have) hold the best so far
  bool dirty = false;  // ... unsorted / not yet cut to top_k
  auto sort_and_cut = [&]() {
    const int m = next_pow2(have < 2 ? 2 : have);
    __syncthreads();
    for (int i = have + tid; i < m; i += kMergeThreads) keys[i] = 0ull;
    __syncthreads();
    block_bitonic_sort_desc(keys, m);
    have = have < M.top_k ? have : M.top_k;
    dirty = false;
    __syncthreads();
  };
  // the list counts are read kMergeThreads at a time (most lists of a query whose work items were not split are empty);
  // non-empty lists are appended while they fit, the buffer sorted and cut to top_k when the next one does not
  for (int l0 = 0; l0 < M.n_lists; l0 += kMergeThreads) {
    __syncthreads();
    if (tid == 0) s_nnz = 0;
    __syncthreads();
    const int l = l0 + tid;
    const int c = l < M.n_lists ? M.slice_cnt[(size_t)q * M.n_lists + l] : 0;
    if (c > 0) { const int p = atomicAdd(&s_nnz, 1); s_nz[p] = l; s_cnt[p] = c; }
    __syncthreads();
    const int nnz = s_nnz;
    // exclusive prefix of the non-empty lists' counts (block-wide Hillis-Steele scan): the keys of the chunk become one flat
    // range that the 256 threads read with independent loads, instead of list after list
    s_off[tid] = tid < nnz ? s_cnt[tid] : 0;
    __syncthreads();
    for (int d = 1; d < kMergeThreads; d <<= 1) {
      const int v = tid >= d ? s_off[tid - d] : 0;
      __syncthreads();
      s_off[tid] += v;
      __syncthreads();
    }
    // (inclusive scan: s_off[i] = keys of lists 0..i of the chunk)
    int i0 = 0;
    while (i0 < nnz) {
      // the longest run of lists [i0, i1) that fits behind the keys kept so far (worst case: no key below the floor)
      const int base = i0 ? s_off[i0 - 1] : 0;
      if (have + (s_off[i0] - base) > kMergeCap) sort_and_cut();   // room for list i0 at least (after the cut: <= top_k + top_k keys)
      int i1 = i0 + 1;
      while (i1 < nnz && have + (s_off[i1] - base) <= kMergeCap) ++i1;
      const int n_flat = s_off[i1 - 1] - base;
      __syncthreads();
      if (tid == 0) s_fill = have;
      __syncthreads();
      for (int e0 = tid; e0 < n_flat; e0 += 4 * kMergeThreads) {   // four independent loads in flight per thread
        uint64_t kk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * kMergeThreads;
          kk[u] = 0ull;
          if (e < n_flat) {
            int lo = i0, hi = i1;   // list of flat element e: the first i in [i0, i1) with s_off[i] - base > e
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_off[mid] - base > e) hi = mid; else lo = mid + 1; }
            kk[u] = M.slice_keys[((size_t)q * M.n_lists + s_nz[lo]) * M.top_k + (e - (lo > i0 ? s_off[lo - 1] - base : 0))];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (kk[u] != 0ull && kk[u] >= floor_key) keys[atomicAdd(&s_fill, 1)] = kk[u];   // (a real key is never 0)
      }
      __syncthreads();
      if (s_fill > have) dirty = true;
      have = s_fill;
      i0 = i1;
    }
  }
  if (dirty) sort_and_cut();
  if (M.n_lists <= 0) have = 0;
  __syncthreads();
  for (int i = tid; i < have; i += kMergeThreads) {
    uint64_t k = keys[i];
    M.out_docs[(size_t)q * M.top_k + i] = key_doc(k) + M.doc_base;
    M.out_scores[(size_t)q * M.top_k + i] = key_score(k);
  }
  if (tid == 0) {
    M.out_counts[q] = have;
    if (M.terminate_after > 0 && M.terminated && M.total_hits && (long long)M.total_hits[q] > M.terminate_after) M.terminated[q] = 1;
    const bool term = M.terminated && M.terminated[q];
    if (M.out_total) {
      long long t = M.total_hits ? (long long)M.total_hits[q] : 0ll;
      if (M.known_hits && M.pruned && M.pruned[q] && (long long)M.known_hits[q] > t) t = (long long)M.known_hits[q];   // a lower bound either way
      M.out_total[q] = t;
    }
    if (M.out_flags) M.out_flags[q] = ((M.pruned && M.pruned[q]) || term ? 1 : 0) | (term ? 2 : 0);
  }
}

// TopDocs.merge over lists of (doc, score) pairs (cross-shard merge after the NCCL all-gather).
struct MergePairsLaunch {
  const int32_t* docs; const float* scores; const int32_t* counts;  // list l: docs + l * stride_hits [nq][top_k], counts + l * stride_counts [nq]
  int64_t stride_hits, stride_counts;   // elements between consecutive lists
  int32_t n_lists, top_k, nq;
  int32_t* out_docs; float* out_scores; int32_t* out_counts;
  // optional (packed records): totalHits summed, flags ORed over the shards (TopDocs.merge: relation GTE if any input is)
  const long long* totals; const int32_t* flags; int64_t stride_totals, stride_flags;
  long long* out_total; int32_t* out_flags;
};

__global__ void __launch_bounds__(kMergeThreads) merge_pairs_kernel(MergePairsLaunch M) {
  __shared__ uint64_t keys[kMergeCap];
  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  int have = 0;
  int l = 0;
  while (l < M.n_lists) {
    int fill = have;
    int l_end = l;
    __syncthreads();
    while (l_end < M.n_lists) {
      int c = M.counts[(size_t)l_end * M.stride_counts + q];
      if (fill + c > kMergeCap) break;
      size_t base = (size_t)l_end * M.stride_hits + (size_t)q * M.top_k;
      for (int i = tid; i < c; i += kMergeThreads) keys[fill + i] = make_key(M.scores[base + i], M.docs[base + i]);
      fill += c; ++l_end;
    }
    l = l_end;
    int m = next_pow2(fill < 2 ? 2 : fill);
    for (int i = fill + tid; i < m; i += kMergeThreads) keys[i] = 0ull;
    __syncthreads();
    block_bitonic_sort_desc(keys, m);
    have = fill < M.top_k ? fill : M.top_k;
  }
  __syncthreads();
  for (int i = tid; i < have; i += kMergeThreads) {
    uint64_t k = keys[i];
    M.out_docs[(size_t)q * M.top_k + i] = key_doc(k);
    M.out_scores[(size_t)q * M.top_k + i] = key_score(k);
  }
  if (tid == 0) {
    M.out_counts[q] = have;
    if (M.out_total) {
      long long t = 0; int32_t f = 0;
      for (int l2 = 0; l2 < M.n_lists; ++l2) { t += M.totals[(size_t)l2 * M.stride_totals + q]; f |= M.flags[(size_t)l2 * M.stride_flags + q]; }
      M.out_total[q] = t; M.out_flags[q] = f;
    }
  }
}

}  // namespace nrtgpu
