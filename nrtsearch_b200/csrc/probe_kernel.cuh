// posting_probe_kernel -- the batched BooleanQuery engine of round 2 for queries of <= 4 term clauses that a posting
// list can lead. It replaces, for a whole batch of queries, what Lucene does per query in
//   MaxScoreBulkScorer (pure disjunctions, essential / non-essential partition),
//   ConjunctionDISI / BlockMaxConjunctionBulkScorer (the rarest required list leads, the others are advanced to it),
//   ReqExclBulkScorer / ReqOptSumScorer (MUST_NOT and optional clauses looked up per candidate)
// behind IndexSearcher.search (reference src/main/java/com/yelp/nrtsearch/server/handler/SearchHandler.java:1412).
//
// Design (B200-first, nothing like the per-document iterator chain of the reference):
//   * work item = (query, doc slice, part), claimed from an atomic queue by PERSISTENT CTAs (3 or 4 per SM): no per-CTA
//     launch cost; slice-major so that the CTAs resident together probe the same doc range of the dense tf planes in
//     L2; a heavy (query, slice) is split into 2..16 parts so that no item is a large share of the launch; warm-up
//     items first: a query sweeps the first 32K postings of its highest-bound list over the whole shard (lower-bound
//     scores, nothing output) and publishes a threshold before any of its other items runs;
//   * the kernel is data parallel over the DRIVER postings of the item: every thread takes postings of the lists
//     that lead (the essential lists of a disjunction, the rarest required list of a conjunction), and PROBES every
//     other list for the doc: a byte gather from the list's dense tf plane (index-time direct-address bytes, L2), or
//     a granule-narrowed binary search of the list's slice segment staged in shared memory by 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx). No window array, no scatter, no per-window barriers: one barrier per
//     round of kR * 256 postings, whose gathers are all in flight together;
//   * pure disjunctions: rank-safe tf-pattern bound test before a doc is appended, UNSCORED, to the candidate
//     buffer; the exact Lucene floats (BM25Scorer expression, double clause sums) are computed at the buffer flush;
//     MAXSCORE partition from index-time list bounds and the query's running threshold theta (one global 64-bit word
//     per query, atomicMax: the device analogue of LazyMaxScoreAccumulator.java:21-70);
//   * exact totalHits without sweeping the densest list: hits = |L1| + sum over the other lists of the postings whose
//     doc is in no earlier list (inclusion by ownership), so in ScoreMode.COMPLETE a non-essential dense list with a
//     plane contributes its posting count and is never read;
//   * anything else (MUST / FILTER / MUST_NOT, ranges, several fields, deletes, minimumNumberShouldMatch): the
//     generic instantiation evaluates the clause tree per surviving driver posting (leap-frog: a doc missing a
//     required list is dropped after the probes, before any norm / doc-value gather).
// Results are bit-identical to the exhaustive oracle (tests/test_gpu_parity.py, tests/test_gpu_probe.py).
#pragma once
#include "stream_kernel.cuh"
#include "sort_kernel.cuh"
#include "collect_kernel.cuh"

namespace nrtgpu {
namespace v3 {

using v2::bulk_g2s;
using v2::mbar_arrive_expect_tx;
using v2::mbar_init;
using v2::mbar_try_wait;
using v2::smem_u32;

// profiling builds only (-DNRT_PROBE_KNOCK): parts of the kernel can be disabled at run time through ProbeLaunch::knock
#ifdef NRT_PROBE_KNOCK
#define NRT_KNOCK(bit) ((L.knock & (bit)) != 0)
#else
#define NRT_KNOCK(bit) false
#endif
constexpr int kT = 4;
// Two launch configurations of the same kernel: 3 CTAs / SM with an 8192-posting stage (80 registers; best for the
// MAXSCORE-pruned sweeps of TOP_SCORES, whose sparse leading lists want the larger stage) and 4 CTAs / SM with a
// 6656-posting stage (64 registers; best where every posting is visited: ScoreMode.COMPLETE and the generic clause
// evaluation, both latency bound on their gathers: 32 resident warps hide more of it than 24).
constexpr int kCtasA = 3, kStageA = 8192;
constexpr int kCtasB = 4, kStageB = 6656;
constexpr int kThreads = 256;
constexpr int kLogGran = v2::kLogGran;       // 1024-doc granules: the granularity of the index-time skip data (gran_tab)
constexpr int kGran = 1 << kLogGran;
constexpr int kMaxSliceGran = 512;           // a slice spans at most 512K docs (its granule offsets live in shared memory)
constexpr int kAlign = 16;                   // staged segments start on 16-posting boundaries (TMA: 16-byte aligned tf bytes)
constexpr int kLongReserve = kT * (kGran + 2 * kAlign);   // one granule of every long list always fits
#ifndef NRT_PROBE_R
#define NRT_PROBE_R 2
#endif
constexpr int kR = NRT_PROBE_R;              // driver postings per thread per round (their gathers are in flight together)
constexpr int kCand = 1024;                  // candidate buffer entries
constexpr int kMaxTopK = kCand / 2;
constexpr int kUbt = 6 * 6 * 6 * 6;
constexpr int kWarmGran = v2::kWarmGran;
constexpr uint32_t kPiece = 8192;            // bytes per bulk copy
constexpr uint32_t kTfInexact = 0xFEu;       // tf byte of a plane probe whose 2-bit code saturated (tf >= 3): the exact byte is
                                             // fetched from the byte plane when the doc is scored (rare); >= 5 for the bound table

enum { kAbsent = 0, kLong = 1, kShort = 2, kPlane = 3, kGlobal = 4 };

struct ProbeLaunch {
  DevIndexView ix;
  const DevClause* clauses;
  const DevQuery* queries;
  const int32_t* work_query;
  const int32_t* work_slice;     // slice | part << 16 | log2(parts) << 20 | flags << 24 (4: sweep warm-up item, see the kernel; 1: warm-up item = first kWarmGran granules of slice 0, 2: slice-0 item behind them)
  const uint32_t* sbounds;       // [nq][kT][n_slices * parts_max + 2]: postings of the slot's list below every part boundary, the shard end, the warm-up boundary
  const uint8_t* field_min_norm;
  unsigned int* work_counter;    // queue head
  unsigned long long* stats;     // optional [8]: items, item cycles, runs, driver postings, flushes, staged postings, set-up cycles, rounds
  int32_t n_work, n_lists, n_slices, top_k;
  int32_t parts_max;             // result lists / boundary entries per slice (a heavy (query, slice) is split into up to this many items)
  int32_t slice_docs;            // multiple of kGran, <= kMaxSliceGran * kGran
  int32_t n_gran;
  int64_t threshold;             // INT32_MAX: ScoreMode.COMPLETE (exact counts)
  int32_t* pruned;
  uint64_t* theta;
  unsigned long long* total_hits;
  uint64_t* slice_keys;
  int32_t* slice_cnt;
  // deadline (SearchCutoffWrapper.java:164-174, checked at work-item boundaries = the reference's per-segment check):
  // the first work item of the run stamps clock0 with %globaltimer; an item claimed more than deadline_ns later is
  // skipped and its query flagged. terminateAfter (TerminateAfterWrapper.java:150-162): a query that has already
  // collected that many hits takes no further work items.
  long long deadline_ns;         // 0: no deadline; < 0: already expired
  unsigned long long* clock0;
  int32_t* timed_out;            // [nq]
  long long terminate_after;     // 0: none
  int32_t* terminated;           // [nq]
  // sort-by-field (generic instantiation): the key of a hit is (order-preserving code of its sort value, ~doc) instead
  // of (score, ~doc); see sort_kernel.cuh
  int32_t sort_kind, sort_reverse;
  const uint32_t* sort_codes;    // [n_docs] codes of the sort column (0 = doc without a value)
  const uint32_t* sort_missing_code;   // [1] code of the sort's missing value
  const AggLaunch* aggs;         // additional collectors (generic instantiation; device pointer, NULL: none)
  const unsigned long long* known_hits;   // optional [nq]: docs KNOWN to match (the longest list of a pure disjunction on a shard without
                                          // deletes): lets pruning start before that many hits were collected (totalHits > threshold is a fact)
  int32_t knock;                 // profiling only (NRTGPU_KNOCK): 1 no plane gathers, 2 no searches, 4 no appends, 8 no sweep
};

template <int kStageT>
struct alignas(128) ProbeSmemT {
  static constexpr int kStage = kStageT;             // postings of the searched lists resident in shared memory (5 B each)
  static constexpr int kShortMax = kStageT - kLongReserve;   // lists without skip data are staged whole, up to this many postings
  static_assert(kShortMax >= 1024, "stage too small");
  int32_t sdocs[kStageT];
  uint8_t sf8[kStageT];
  uint32_t gb[kT][kMaxSliceGran + 4];   // granule offsets of the slice for lists with skip data (relative to the list's first posting)
  uint64_t cand[kCand];
  float ubt[kUbt];
  float uval[kT][4];
  DevClause cl[kMaxClauses];
  DevQuery q;
  uint64_t stage_bar;
  // per slot (CTA-uniform, written by thread 0 / threads < kT between barriers)
  const int32_t* s_gdocs[kT];      // global postings of the list
  const uint8_t* s_gf8[kT];
  const uint8_t* s_plane[kT];      // byte plane (exact min(tf, 255) per doc) of the list, or NULL
  const uint8_t* s_plane2[kT];     // 2-bit plane (min(tf, 3), four docs per byte): what the probes gather
  float s_weight[kT];
  float s_ub[kT];
  int32_t s_kind[kT];
  int32_t s_clause[kT];
  int32_t s_field[kT];
  uint32_t s_ia[kT], s_ib[kT];     // item bounds (postings relative to the list's first)
  uint32_t s_ra[kT], s_rb[kT];     // run bounds
  int32_t s_sdelta[kT];            // staged lists: smem index of posting x = x + s_sdelta
  uint32_t s_pbm[kT];              // post_base mod 16 (alignment of the list inside the global posting arrays)
  int32_t s_row[kT];               // row of the index-time granule offsets (gb[] holds the slice's part), -1: none
  uint32_t s_need[kT];             // slots a driver posting of this slot probes
  uint32_t s_candbelow[kT];        // word bytes of the lists that own a doc before this slot (candidate emission)
  uint32_t s_cntbefore[kT];        // ... (hit counting)
  uint32_t s_pre[kT + 1];          // prefix of the driver postings of the run
  uint32_t drv_mask, ess_mask, plane_mask, long_mask, short_mask, global_mask;
  int32_t short_total;             // staged postings of the short lists (aligned)
  int32_t g1;                      // end of the current run (granule of the slice)
  int32_t run_d0, run_d1;          // doc range of the current run
  int32_t staged;                  // the current run issued TMA copies
  int wi;
  int skip;                        // the claimed item is not processed (abort flag set)
  int cand_count;
  int n_keys;
  unsigned long long hits0;
  unsigned long long hits_known;   // max(hits0, docs known to match)
  int theta_dec;                   // 1: the item publishes (k-th key - 1) as threshold (sweep warm-up: its candidates are not output)
  unsigned long long theta;
};
static_assert(sizeof(ProbeSmemT<kStageA>) <= 232448 / kCtasA - 1024 && sizeof(ProbeSmemT<kStageB>) <= 232448 / kCtasB - 1024, "ProbeSmem exceeds the per-CTA shared memory budget");

// A staged segment [a, b) of a list (postings relative to the list's first) is copied from the enclosing 16-posting
// aligned range of the GLOBAL posting arrays (TMA needs 16-byte aligned tf bytes): with pbm = post_base mod 16 the
// copy starts at list-relative posting seg_first(a, pbm) (may be negative: the tail of the previous list) and holds
// seg_n(a, b, pbm) postings.
__device__ __forceinline__ int32_t seg_first(uint32_t a, uint32_t pbm) { return (int32_t)((a + pbm) & ~(uint32_t)(kAlign - 1)) - (int32_t)pbm; }
__device__ __forceinline__ uint32_t seg_n(uint32_t a, uint32_t b, uint32_t pbm) {
  return b > a ? ((b + pbm + kAlign - 1) & ~(uint32_t)(kAlign - 1)) - ((a + pbm) & ~(uint32_t)(kAlign - 1)) : 0u;
}

// Universal clause evaluation of one doc given the tf word of its term slots (Lucene BooleanScorerSupplier semantics,
// as v2::evaluate_doc_generic: conjunction / disjunction sums in double, ReqOptSumScorer float add when msm == 0).
template <typename SM>
__device__ __noinline__ bool evaluate_doc(const ProbeLaunch& L, const SM& sm, int32_t doc, uint32_t word, float* out_score) {
  const DevQuery& q = sm.q;
  const uint32_t m = v2::presence4(word);
  if ((m & q.req_term_mask) != q.req_term_mask) return false;
  if (m & q.not_term_mask) return false;
  if (L.ix.live_bits && !((L.ix.live_bits[doc >> 5] >> (doc & 31)) & 1u)) return false;
  // doc-value clauses first: a doc that fails a required range (or hits an excluded one) is dropped before any norm is
  // gathered or score computed (what ConjunctionDISI does by advancing the cheapest iterators first)
  uint32_t range_present = 0;
  if (q.has_nonterm)
    for (int i = 0; i < q.n_clauses; ++i) {
      const DevClause& c = sm.cl[i];
      if (c.kind != NRTGPU_RANGE_I64) continue;
      const bool p = range_matches(L.ix, c.col, doc, c.lo, c.hi);
      if (p) { if (c.occur == NRTGPU_MUST_NOT) return false; range_present |= 1u << i; }
      else if (c.occur == NRTGPU_MUST || c.occur == NRTGPU_FILTER) return false;
    }
  double must_sum = 0.0, should_sum = 0.0;
  int n_should = 0;
  int cur_field = -1;
  uint32_t nb = 1u;
  for (int i = 0; i < q.n_clauses; ++i) {
    const DevClause& c = sm.cl[i];
    bool present;
    float s = 0.0f;
    if (c.kind == NRTGPU_TERM) {
      uint32_t b = (word >> (8 * c.slot)) & 0xffu;
      present = b != 0;
      if (present && c.scoring) {
        if (b == kTfInexact && sm.s_plane[c.slot]) b = (uint32_t)__ldg(sm.s_plane[c.slot] + doc);   // saturated 2-bit code
        if (c.field != cur_field) {
          cur_field = c.field;
          const uint8_t* nrm = L.ix.norms[c.field];
          nb = nrm ? (uint32_t)__ldg(nrm + doc) : 1u;
        }
        const float f = (b == 255u) ? exact_freq_slow<uint32_t>(L.ix, c, doc) : (float)b;
        s = bm25_score(c.weight, f, __ldg(&L.ix.caches[c.field * 256 + nb]));
      }
    } else if (c.kind == NRTGPU_RANGE_I64) {
      present = (range_present >> i) & 1u;
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
      default: return false;
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

// exact score of a doc of a pure single-field disjunction: double sum, in slot (= clause) order, of Lucene's BM25 float
// expression for the slots present (BM25Scorer.score; DisjunctionSumScorer / MaxScoreBulkScorer sum in double)
template <typename SM>
__device__ __forceinline__ float score_disjunction(const ProbeLaunch& L, const SM& sm, const uint8_t* norms0, int n_term,
                                                   int32_t doc, uint32_t word) {
  const uint32_t nb = norms0 ? (uint32_t)__ldg(norms0 + doc) : 1u;
  double sum = 0.0;
#pragma unroll
  for (int s = 0; s < kT; ++s) {
    if (s >= n_term) break;
    uint32_t b = (word >> (8 * s)) & 0xffu;
    if (b == 0) continue;
    if (b == kTfInexact && sm.s_plane[s]) b = (uint32_t)__ldg(sm.s_plane[s] + doc);   // saturated 2-bit code: the exact byte
    const float f = (b == 255u) ? exact_freq_slow<uint32_t>(L.ix, sm.cl[sm.s_clause[s]], doc) : (float)b;
    sum += (double)bm25_score(sm.s_weight[s], f, __ldg(&L.ix.caches[sm.s_field[s] * 256 + nb]));
  }
  return (float)sum;
}

// Candidate buffer flush (all threads). Entries [0, n_keys) are keys kept by the previous flush; the rest are keys
// (generic) or unscored (tf word << 32 | doc) pairs (pure disjunctions) which are scored here, one per thread, so the norm
// loads of the whole buffer overlap. Keeps the best top_k, publishes the k-th key as the query's threshold.
template <bool kSimple, typename SM>
__device__ __noinline__ void flush_candidates(const ProbeLaunch& L, SM& sm, const uint8_t* norms0, int n_term,
                                                 bool has_after, uint64_t after_key, int top_k, uint64_t* g_theta) {
  __syncthreads();
  int n = sm.cand_count;
  if (n > kCand) n = kCand;
  if (kSimple) {
    const unsigned long long theta = sm.theta;
    const int n_keys = sm.n_keys;
    constexpr int kPer = kCand / kThreads;
    uint64_t mine[kPer];
    // every gather of the thread's candidates is issued before the first score is computed: the norm byte, and the
    // exact tf byte of every slot whose 2-bit plane code saturated
    uint32_t nbv[kPer], wv[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int i = n_keys + (int)threadIdx.x + j * kThreads;
      mine[j] = (i < n) ? sm.cand[i] : 0ull;
      const int32_t doc = (int32_t)(uint32_t)mine[j];
      nbv[j] = (mine[j] && norms0) ? (uint32_t)__ldg(norms0 + doc) : 1u;
      uint32_t w = (uint32_t)(mine[j] >> 32);
#pragma unroll
      for (int s2 = 0; s2 < kT; ++s2)
        if (((w >> (8 * s2)) & 0xffu) == kTfInexact && sm.s_plane[s2])
          w = (w & ~(0xffu << (8 * s2))) | ((uint32_t)__ldg(sm.s_plane[s2] + doc) << (8 * s2));
      wv[j] = w;
    }
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      uint64_t key = 0ull;
      if (mine[j]) {
        const int32_t doc = (int32_t)(uint32_t)mine[j];
        double sum = 0.0;   // BM25Scorer.score per slot (float), DisjunctionSumScorer / MaxScoreBulkScorer sum in double, slot order
#pragma unroll
        for (int s2 = 0; s2 < kT; ++s2) {
          const uint32_t b = (wv[j] >> (8 * s2)) & 0xffu;
          if (s2 >= n_term || b == 0u) continue;
          const float f = (b == 255u) ? exact_freq_slow<uint32_t>(L.ix, sm.cl[sm.s_clause[s2]], doc) : (float)b;
          sum += (double)bm25_score(sm.s_weight[s2], f, __ldg(&L.ix.caches[sm.s_field[s2] * 256 + nbv[j]]));
        }
        key = make_key((float)sum, doc);
        if (!(key > theta) || (has_after && !(key < after_key))) key = 0ull;   // a real key is never 0
      }
      mine[j] = key;
    }
    __syncthreads();
    if (threadIdx.x == 0) sm.cand_count = n_keys;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPer; ++j)
      if (mine[j]) sm.cand[atomicAdd(&sm.cand_count, 1)] = mine[j];
    __syncthreads();
    n = sm.cand_count;
    if (n < top_k) {   // fewer than top_k keys in all: nothing to drop, no k-th key to publish (the slice merge sorts)
      __syncthreads();   // every thread has read cand_count before anybody appends again
      if (threadIdx.x == 0) {
        sm.n_keys = n;
        const unsigned long long g = *(volatile unsigned long long*)g_theta;
        if (g > sm.theta) sm.theta = g;
      }
      __syncthreads();
      return;
    }
  }
  const int m = next_pow2(n < 2 ? 2 : n);
  const long long ts = L.stats ? clock64() : 0ll;
  for (int i = n + threadIdx.x; i < m; i += kThreads) sm.cand[i] = 0ull;
  __syncthreads();
  block_bitonic_sort_desc(sm.cand, m);
  if (L.stats && threadIdx.x == 0) atomicAdd(&L.stats[15], (unsigned long long)(clock64() - ts));
  if (threadIdx.x == 0) {
    const int keep = n < top_k ? n : top_k;
    sm.cand_count = keep;
    sm.n_keys = keep;
    if (keep == top_k) {
      const unsigned long long kth = sm.cand[top_k - 1] - (unsigned long long)sm.theta_dec;
      const unsigned long long old = atomicMax((unsigned long long*)g_theta, kth);
      const unsigned long long t = old > kth ? old : kth;
      if (t > sm.theta) sm.theta = t;
    } else {
      const unsigned long long g = *(volatile unsigned long long*)g_theta;
      if (g > sm.theta) sm.theta = g;
    }
  }
  __syncthreads();
}

// binary search of doc in the sorted smem range [l, h); returns the tf byte (0 = absent)
template <typename SM>
__device__ __forceinline__ uint32_t probe_smem(const SM& sm, int l, int h, int32_t doc) {
  const int end = h;
  while (l < h) {
    const int mid = (l + h) >> 1;
    if (sm.sdocs[mid] < doc) l = mid + 1; else h = mid;
  }
  return (l < end && sm.sdocs[l] == doc) ? (uint32_t)sm.sf8[l] : 0u;
}
__device__ __noinline__ uint32_t probe_global(const int32_t* docs, const uint8_t* f8, uint32_t l, uint32_t h, int32_t doc) {
  const uint32_t end = h;
  while (l < h) {
    const uint32_t mid = (l + h) >> 1;
    if (__ldg(docs + mid) < doc) l = mid + 1; else h = mid;
  }
  return (l < end && __ldg(docs + l) == doc) ? (uint32_t)__ldg(f8 + l) : 0u;
}

// kStats: the profiling instantiation (NRTGPU_DEBUG_MODES=1) keeps cycle counters; the production one has none of their registers
template <bool kSimple, bool kStats, int kCtas, int kStageT>
__global__ void __launch_bounds__(kThreads, kCtas) posting_probe_kernel(const __grid_constant__ ProbeLaunch L) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using ProbeSmem = ProbeSmemT<kStageT>;
  constexpr int kStage = kStageT, kShortMax = ProbeSmem::kShortMax;
  ProbeSmem& sm = *reinterpret_cast<ProbeSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  if (tid == 0) {
    mbar_init(&sm.stage_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  uint32_t stage_parity = 0;   // phase of stage_bar the next staged run completes (tracked identically by every thread)
  const int gran_per_slice = L.slice_docs >> kLogGran;
  const int fine = (gran_per_slice + L.parts_max - 1) / L.parts_max;   // granules per finest part of a slice
  const int sb_stride = L.n_slices * L.parts_max + 2;
  const long long t_cta = kStats ? clock64() : 0ll;

  for (;;) {
    __syncthreads();   // the previous item is retired (also orders the mbarrier init before its first use)
    if (tid == 0) {
      const int w = (int)atomicAdd(L.work_counter, 1u);
      sm.wi = w;
      sm.skip = 0;
      if (L.deadline_ns && w < L.n_work) {
        bool late = L.deadline_ns < 0;   // the request's budget was spent before the launch
        if (!late) {
          unsigned long long now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          unsigned long long t0 = atomicCAS(L.clock0, 0ull, now);
          if (t0 == 0ull) t0 = now;
          late = now > t0 && now - t0 > (unsigned long long)L.deadline_ns;   // (another CTA may have stamped clock0 after this one read the timer)
        }
        if (late) { sm.skip = 1; L.timed_out[L.work_query[w]] = 1; }   // drain the queue
      }
    }
    __syncthreads();
    const int wi = sm.wi;
    if (wi >= L.n_work) break;
    if (sm.skip) continue;
    const long long t_start = kStats ? clock64() : 0ll;
    const int qi = L.work_query[wi];
    const int slice_raw = L.work_slice[wi];
    const int slice = slice_raw & 0xffff;
    const int wflags = slice_raw >> 24;
    // Sweep warm-up item (flags 4): the first 32K postings of the query's highest-bound list, over the WHOLE shard, probing
    // only the lists with tf planes (the other lists count as absent: scores are lower bounds). Nothing is output; the
    // k-th best lower-bound key minus one becomes the query's threshold before any other item of the query runs --
    // the docs that hold the query's rarest term are where its top-k is, a far better sample than the first 32K docs.
    const bool sweep_warm = (wflags & 4) != 0;
    const int warm_slot = (slice_raw >> 16) & 0xf;
    const int part = sweep_warm ? 0 : ((slice_raw >> 16) & 0xf), lparts = (slice_raw >> 20) & 0xf;   // part `part` of 2^lparts of the slice
    const int ncl = L.queries[qi].n_clauses, cbeg = L.queries[qi].clause_begin, n_term = L.queries[qi].n_term;
    if (tid == 0) {
      sm.q = L.queries[qi];
      sm.cand_count = 0;
      sm.n_keys = 0;
      sm.theta = *(volatile unsigned long long*)&L.theta[qi];
      sm.hits0 = *(volatile unsigned long long*)&L.total_hits[qi];
      { const unsigned long long kn = L.known_hits ? L.known_hits[qi] : 0ull; sm.hits_known = kn > sm.hits0 ? kn : sm.hits0; }
      sm.theta_dec = sweep_warm ? 1 : 0;
    }
    if (tid < ncl) sm.cl[tid] = L.clauses[cbeg + tid];
    if (tid >= 32 && tid < 32 + kT) { const int s = tid - 32; sm.s_kind[s] = kAbsent; sm.s_ia[s] = 0; sm.s_ib[s] = 0; sm.s_ra[s] = 0; sm.s_rb[s] = 0;
                                      sm.s_plane[s] = nullptr; sm.s_plane2[s] = nullptr; sm.s_gdocs[s] = nullptr; sm.s_gf8[s] = nullptr; sm.s_weight[s] = 0.f; sm.s_ub[s] = 0.f;
                                      sm.s_clause[s] = 0; sm.s_field[s] = 0; sm.s_sdelta[s] = 0; sm.s_pbm[s] = 0; sm.s_row[s] = -1; }
    const int g_first = slice * gran_per_slice;
    const int g_count = min(gran_per_slice, L.n_gran - g_first);
    // granule range of the item inside its slice, and the entries of the boundary table that hold its posting bounds
    const int kfine = L.parts_max >> lparts;   // finest parts per part of this item
    int g_lo = min(g_count, part * kfine * fine);
    int g_hi = ((part + 1) * kfine >= L.parts_max) ? g_count : min(g_count, (part + 1) * kfine * fine);
    const int e_lo = slice * L.parts_max + part * kfine;
    int e_hi = slice * L.parts_max + (part + 1) * kfine;   // (part + 1) * kfine == parts_max: entry 0 of the next slice / the end entry
    const int e_warm = L.n_slices * L.parts_max + 1;
    if (wflags & 2) g_lo = max(g_lo, min(g_count, kWarmGran));
    if (wflags & 1) { g_hi = min(g_count, kWarmGran); e_hi = e_warm; }
    const int e_lo2 = sweep_warm ? 0 : e_lo;                              // whole-shard bounds of every list; one dummy run of one granule
    if (sweep_warm) { g_lo = 0; g_hi = 1; e_hi = L.n_slices * L.parts_max; }
    __syncthreads();   // B1: query + clauses resident
    // terminateAfter (TerminateAfterWrapper.java:150-162): a query that has collected enough hits stops collecting
    if (L.terminate_after > 0 && (long long)sm.hits0 >= L.terminate_after) {
      if (tid == 0) L.terminated[qi] = 1;
      continue;
    }
    // ---- per-slot descriptors (one thread per clause), granule offsets of the lists with skip data (all threads)
    if (tid < ncl && sm.cl[tid].kind == NRTGPU_TERM) {
      const DevClause& c = sm.cl[tid];
      const int s = c.slot;
      const uint32_t* sb = L.sbounds + ((size_t)qi * kT + s) * sb_stride;
      uint32_t a = sb[e_lo2];
      uint32_t b = sb[e_hi];
      if (wflags & 2) a = max(a, sb[e_warm]);
      if (sweep_warm && s == warm_slot) b = min(b, a + 32768u);
      sm.s_ia[s] = a; sm.s_ib[s] = max(a, b);
      sm.s_gdocs[s] = L.ix.post_docs + c.post_base;
      sm.s_gf8[s] = L.ix.post_f8 + c.post_base;
      const bool has_plane = c.plane >= 0 && L.ix.dense_tf != nullptr && L.ix.dense_tf2 != nullptr;
      sm.s_plane[s] = has_plane ? L.ix.dense_tf + (size_t)c.plane * (size_t)L.ix.dense_stride : nullptr;
      sm.s_plane2[s] = has_plane ? L.ix.dense_tf2 + (size_t)c.plane * (size_t)(L.ix.dense_stride >> 2) : nullptr;
      sm.s_kind[s] = has_plane ? kPlane : (c.gran_row >= 0 ? kLong : kShort);   // kShort may become kGlobal below
      if (sweep_warm && !has_plane && s != warm_slot) sm.s_kind[s] = kAbsent;   // not probed by the sweep warm-up: counts as absent
      sm.s_weight[s] = c.weight; sm.s_ub[s] = c.ub; sm.s_clause[s] = tid; sm.s_field[s] = c.field;
      sm.s_pbm[s] = (uint32_t)(c.post_base & (int64_t)(kAlign - 1)); sm.s_row[s] = sweep_warm ? -1 : c.gran_row;
    }
    for (int i = 0; i < ncl; ++i) {
      const DevClause& c = sm.cl[i];
      if (c.kind != NRTGPU_TERM || c.gran_row < 0 || sweep_warm) continue;
      const uint32_t* row = L.ix.gran_tab + (size_t)c.gran_row * (size_t)(L.n_gran + 1) + g_first;
      for (int g = g_lo + tid; g <= g_hi; g += kThreads) sm.gb[c.slot][g] = __ldg(row + g);
    }
    if (kSimple && tid >= 64 && tid < 64 + 4 * kT) {   // per-slot score bounds at tf = 1..4 (shortest field length present)
      const int s = (tid - 64) >> 2, c = ((tid - 64) & 3) + 1;
      float u = 0.0f;
      for (int i = 0; i < ncl; ++i)
        if (sm.cl[i].kind == NRTGPU_TERM && sm.cl[i].slot == s) {
          const uint32_t nbmin = (sm.q.single_field >= 0 && L.field_min_norm) ? (uint32_t)L.field_min_norm[sm.q.single_field] : 0u;
          u = bm25_score(sm.cl[i].weight, (float)c, __ldg(&L.ix.caches[sm.cl[i].field * 256 + nbmin]));
        }
      sm.uval[s][c - 1] = u;
    }
    __syncthreads();   // B2: descriptors, granule offsets, bound values
    if (kSimple) {
      // ubt[sum min(tf_s, 5) * 6^s]: the double clause sum with every term at the shortest field length present
      // (tf >= 5 bounded by the clause weight, the limit tf -> inf) -- an upper bound of the doc's score
      const int n_ubt = n_term >= 4 ? kUbt : (n_term == 3 ? 216 : (n_term == 2 ? 36 : 6));   // patterns of the slots that exist
      for (int i = tid; i < n_ubt; i += kThreads) {
        const int c[kT] = {i % 6, (i / 6) % 6, (i / 36) % 6, i / 216};
        double sum = 0.0;
#pragma unroll
        for (int t = 0; t < kT; ++t) {
          float u = 0.0f;
          if (sm.s_kind[t] != kAbsent && c[t] > 0) u = (c[t] <= 4) ? sm.uval[t][c[t] - 1] : sm.s_weight[t];
          sum += (double)u;
        }
        sm.ubt[i] = (float)sum;
      }
    }
    if (tid == 0) {
      // ---- roles. MAXSCORE split (pure disjunctions): the lists whose list-wide bounds sum (double, ascending) below
      // theta.score are non-essential: they never lead, docs found only in them cannot enter the top-k.
      const uint32_t all = (n_term >= 32) ? 0xffffffffu : ((1u << n_term) - 1u);
      uint32_t ne = 0;
      const bool complete = L.threshold >= (int64_t)INT32_MAX;
      if (kSimple && !sweep_warm && sm.theta != 0ull && (complete || (int64_t)sm.hits_known > L.threshold)) {
        const float theta_s = key_score(sm.theta);
        int ord[kT]; int n = 0;
        for (int s = 0; s < n_term; ++s) ord[n++] = s;
        for (int a = 1; a < n; ++a) { const int x = ord[a]; int b = a - 1; while (b >= 0 && sm.s_ub[ord[b]] > sm.s_ub[x]) { ord[b + 1] = ord[b]; --b; } ord[b + 1] = x; }
        double pre = 0.0;
        for (int a = 0; a < n; ++a) {
          const double s2 = pre + (double)sm.s_ub[ord[a]];
          if (!((float)s2 < theta_s)) break;
          pre = s2; ne |= 1u << ord[a];
        }
      }
      if (ne && !complete) L.pruned[qi] = 1;
      // short lists are staged whole at the first run; what does not fit the reserve is searched in global memory
      int st = 0;
      uint32_t pm = 0, lm = 0, shm = 0, gm = 0;
      for (int s = 0; s < n_term; ++s) {
        const int k = sm.s_kind[s];
        if (k == kPlane) pm |= 1u << s;
        else if (k == kLong) lm |= 1u << s;
        else if (k == kShort) {
          const uint32_t a = sm.s_ia[s], b = sm.s_ib[s];
          const int need = (int)seg_n(a, b, sm.s_pbm[s]);
          if (st + need <= kShortMax) { sm.s_sdelta[s] = st - seg_first(a, sm.s_pbm[s]); st += need; shm |= 1u << s; }
          else { sm.s_kind[s] = kGlobal; gm |= 1u << s; }
        }
      }
      if (sweep_warm) { st = 0; lm = 0; shm = 0; gm = 0; }   // the leading list is read in place, nothing is staged or searched
      sm.short_total = st;
      sm.plane_mask = pm; sm.long_mask = lm; sm.short_mask = shm; sm.global_mask = gm;
      uint32_t drv, ess;
      if (kSimple && sweep_warm) {
        drv = ess = 1u << warm_slot;
        for (int t = 0; t < n_term; ++t) { sm.s_candbelow[t] = 0u; sm.s_cntbefore[t] = 0xffffffffu; sm.s_need[t] = pm & ~(1u << t); }   // (cntbefore: nothing is counted)
      } else if (kSimple) {
        ess = all & ~ne;
        int cnt_first = -1;
        if (complete && ne && !L.ix.live_bits) {   // the densest non-essential list with a plane contributes its posting count unread
          uint32_t best = 0;
          for (int s = 0; s < n_term; ++s)
            if (((ne >> s) & 1u) && sm.s_kind[s] == kPlane && sm.s_ib[s] - sm.s_ia[s] >= best) { best = sm.s_ib[s] - sm.s_ia[s]; cnt_first = s; }
        }
        drv = complete ? (cnt_first >= 0 ? all & ~(1u << cnt_first) : all) : ess;
        for (int t = 0; t < n_term; ++t) {
          uint32_t below_ess = 0, before_cnt = 0;
          for (int s = 0; s < t; ++s) {
            if ((ess >> s) & 1u) below_ess |= 0xffu << (8 * s);
            if (s != cnt_first) before_cnt |= 0xffu << (8 * s);
          }
          if (cnt_first >= 0 && cnt_first != t) before_cnt |= 0xffu << (8 * cnt_first);
          sm.s_candbelow[t] = below_ess;
          sm.s_cntbefore[t] = complete ? before_cnt : below_ess;
          uint32_t need = all & ~(1u << t);
          if (complete && !((ess >> t) & 1u)) {   // a non-essential list is swept only to count: probe the earlier lists
            need = 0;
            for (int s = 0; s < n_term; ++s) if (s != t && (s == cnt_first || s < t)) need |= 1u << s;
          }
          sm.s_need[t] = need;
        }
        if (complete && cnt_first >= 0 && g_lo < g_hi)
          atomicAdd(&L.total_hits[qi], (unsigned long long)(sm.s_ib[cnt_first] - sm.s_ia[cnt_first]));
      } else {
        ess = sm.q.dense_driver ? 0u : (sm.q.driver_mask & all);   // dense: every doc of the slice is visited, all lists are probed
        drv = ess;
        for (int t = 0; t < n_term; ++t) {
          uint32_t below = 0;
          for (int s = 0; s < t; ++s) if ((drv >> s) & 1u) below |= 0xffu << (8 * s);
          sm.s_candbelow[t] = below; sm.s_cntbefore[t] = below;
          sm.s_need[t] = all & ~(1u << t);
        }
      }
      sm.drv_mask = drv; sm.ess_mask = ess;
    }
    __syncthreads();   // B3: roles, staging plan

    const int32_t slice_base = slice * L.slice_docs;
    const bool has_after = sm.q.has_after != 0;
    const uint64_t after_key = sm.q.after_key;
    const uint8_t* norms0 = (kSimple && sm.q.single_field >= 0) ? L.ix.norms[sm.q.single_field] : nullptr;
    const uint32_t drv_mask = sm.drv_mask, ess_mask = sm.ess_mask;
    const uint32_t plane_mask = sm.plane_mask, long_mask = sm.long_mask, short_mask = sm.short_mask, global_mask = sm.global_mask;
    const uint8_t* pl0 = sm.s_plane2[0]; const uint8_t* pl1 = sm.s_plane2[1]; const uint8_t* pl2 = sm.s_plane2[2]; const uint8_t* pl3 = sm.s_plane2[3];
    unsigned int my_hits = 0;
    unsigned long long dbg_post = 0; unsigned int dbg_runs = 0, dbg_rounds = 0, dbg_flush = 0, dbg_staged = 0;
    long long dbg_tflush = 0, dbg_twait = 0;
    const long long t_setup = kStats ? clock64() : 0ll;

    const bool dense = !kSimple && sm.q.dense_driver != 0;
    const uint32_t* live = L.ix.live_bits;
    const uint32_t sort_missing = (!kSimple && L.sort_kind == NRTGPU_SORT_COLUMN) ? *L.sort_missing_code : 0u;
    int g0 = g_lo;
    if (drv_mask == 0u && !dense) g0 = g_hi;   // nothing leads (every list non-essential): the slice cannot contribute
    bool first_run = true;
    while (g0 < g_hi) {
      // ---------------- run = the longest granule range [g0, g1) whose long-list segments fit the stage
      if (tid == 0) {
        { const unsigned long long g = *(volatile unsigned long long*)&L.theta[qi]; if (g > sm.theta) sm.theta = g; }
        const int cap = kStage - sm.short_total;
        auto fits = [&](int g1) {
          int tot = 0;
#pragma unroll
          for (int s = 0; s < kT; ++s)
            if ((long_mask >> s) & 1u) tot += (int)seg_n(sm.gb[s][g0], sm.gb[s][g1], sm.s_pbm[s]);
          return tot <= cap;
        };
        int hi = g_hi, lo = g0 + 1;
        if (long_mask && hi > lo && !fits(hi)) {
          while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (fits(mid)) lo = mid; else hi = mid; }
          hi = lo;
        }
        const int g1 = hi;
        sm.g1 = g1;
        const bool whole = (g0 == g_lo && g1 == g_hi);
        const int32_t d0 = slice_base + (g0 << kLogGran);
        const int64_t d1_64 = (int64_t)slice_base + ((int64_t)g1 << kLogGran);
        const int32_t d1 = d1_64 > (int64_t)L.ix.n_docs ? L.ix.n_docs : (int32_t)d1_64;
        sm.run_d0 = d0; sm.run_d1 = d1;
        // run bounds of every list: skip data where the list has it, else a search by doc (staged short lists: in shared
        // memory once resident -- their first-run bounds are fixed up below; lists read from global memory: there)
        for (int s = 0; s < n_term; ++s) {
          const int k = sm.s_kind[s];
          uint32_t a = sm.s_ia[s], b = sm.s_ib[s];
          if (sm.s_row[s] >= 0) { a = max(a, sm.gb[s][g0]); b = min(b, sm.gb[s][g1]); }
          else if (!whole && b > a) {
            if (k == kShort) {
              if (!first_run) {
                const int base = sm.s_sdelta[s];
                int l = base + (int)a, h = base + (int)b;
                while (l < h) { const int m = (l + h) >> 1; if (sm.sdocs[m] < d0) l = m + 1; else h = m; }
                const uint32_t na = (uint32_t)(l - base);
                h = base + (int)b;
                while (l < h) { const int m = (l + h) >> 1; if (sm.sdocs[m] < d1) l = m + 1; else h = m; }
                a = na; b = (uint32_t)(l - base);
              }
            } else {   // kPlane without skip data, kGlobal
              const int32_t* gd = sm.s_gdocs[s];
              uint32_t l = a, h = b;
              while (l < h) { const uint32_t m = (l + h) >> 1; if (__ldg(gd + m) < d0) l = m + 1; else h = m; }
              const uint32_t na = l;
              h = b;
              while (l < h) { const uint32_t m = (l + h) >> 1; if (__ldg(gd + m) < d1) l = m + 1; else h = m; }
              a = na; b = l;
            }
          }
          if (b < a) b = a;
          sm.s_ra[s] = a; sm.s_rb[s] = b;
        }
        // TMA copies: short lists once per item (first run, whole item segment), long lists per run behind them
        uint32_t total = 0;
        if (first_run)
          for (int s = 0; s < n_term; ++s)
            if ((short_mask >> s) & 1u) total += seg_n(sm.s_ia[s], sm.s_ib[s], sm.s_pbm[s]) * 5u;
        for (int s = 0; s < n_term; ++s)
          if ((long_mask >> s) & 1u) total += seg_n(sm.s_ra[s], sm.s_rb[s], sm.s_pbm[s]) * 5u;
        sm.staged = total != 0u;
        if (total) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of the stage precede the async writes
          mbar_arrive_expect_tx(&sm.stage_bar, total);
          auto stage = [&](int s, uint32_t a, uint32_t b, int at) {   // copy the aligned range around [a, b) to sdocs/sf8[at...]
            const uint32_t pbm = sm.s_pbm[s];
            const int32_t f = seg_first(a, pbm);
            const uint32_t n = seg_n(a, b, pbm);
            sm.s_sdelta[s] = at - f;
            const unsigned char* gd = reinterpret_cast<const unsigned char*>(sm.s_gdocs[s] + f);
            const unsigned char* gf = reinterpret_cast<const unsigned char*>(sm.s_gf8[s] + f);
            unsigned char* dd = reinterpret_cast<unsigned char*>(&sm.sdocs[at]);
            unsigned char* df = reinterpret_cast<unsigned char*>(&sm.sf8[at]);
            for (uint32_t o = 0; o < n * 4u; o += kPiece) bulk_g2s(dd + o, gd + o, min(kPiece, n * 4u - o), &sm.stage_bar);
            for (uint32_t o = 0; o < n; o += kPiece) bulk_g2s(df + o, gf + o, min(kPiece, n - o), &sm.stage_bar);
            return (int)n;
          };
          if (first_run)
            for (int s = 0; s < n_term; ++s)
              if (((short_mask >> s) & 1u) && sm.s_ib[s] > sm.s_ia[s])
                stage(s, sm.s_ia[s], sm.s_ib[s], sm.s_sdelta[s] + seg_first(sm.s_ia[s], sm.s_pbm[s]));
          int st = sm.short_total;
          for (int s = 0; s < n_term; ++s)
            if (((long_mask >> s) & 1u) && sm.s_rb[s] > sm.s_ra[s]) st += stage(s, sm.s_ra[s], sm.s_rb[s], st);
        }
        uint32_t pre = 0;   // prefix of the driver postings
        for (int s = 0; s < kT; ++s) { sm.s_pre[s] = pre; if (s < n_term && ((drv_mask >> s) & 1u)) pre += sm.s_rb[s] - sm.s_ra[s]; }
        sm.s_pre[kT] = pre;
      }
      __syncthreads();   // R1: run plan visible
      const int g1 = sm.g1;
      if (sm.staged) {
        const long long tw = kStats ? clock64() : 0ll;
        while (!mbar_try_wait(&sm.stage_bar, stage_parity)) {}
        stage_parity ^= 1u;
        if (kStats) dbg_twait += clock64() - tw;
      }
      if (first_run && g1 < g_hi && short_mask) {   // multi-run item: narrow the (now resident) short lists to the first run's docs
        __syncthreads();
        if (tid == 0) {
          const int32_t d0 = slice_base + (g0 << kLogGran);
          const int64_t d1_64 = (int64_t)slice_base + ((int64_t)g1 << kLogGran);
          const int32_t d1 = d1_64 > (int64_t)L.ix.n_docs ? L.ix.n_docs : (int32_t)d1_64;
          for (int s = 0; s < n_term; ++s)
            if (((short_mask >> s) & 1u) && sm.s_ib[s] > sm.s_ia[s]) {
              const int base = sm.s_sdelta[s];
              int l = base + (int)sm.s_ia[s], h = base + (int)sm.s_ib[s];
              while (l < h) { const int m = (l + h) >> 1; if (sm.sdocs[m] < d0) l = m + 1; else h = m; }
              const uint32_t na = (uint32_t)(l - base);
              h = base + (int)sm.s_ib[s];
              while (l < h) { const int m = (l + h) >> 1; if (sm.sdocs[m] < d1) l = m + 1; else h = m; }
              sm.s_ra[s] = na; sm.s_rb[s] = (uint32_t)(l - base);
            }
          uint32_t pre = 0;
          for (int s = 0; s < kT; ++s) { sm.s_pre[s] = pre; if (s < n_term && ((drv_mask >> s) & 1u)) pre += sm.s_rb[s] - sm.s_ra[s]; }
          sm.s_pre[kT] = pre;
        }
        __syncthreads();
      }
      if (kStats) ++dbg_runs;
      // ---------------- rounds over the driver postings of the run
      const uint32_t n_total = sm.s_pre[kT];
      if (kStats && tid == 0) { dbg_post += n_total; dbg_staged += sm.staged ? 1u : 0u; }
      // No barrier between rounds: every warp streams through its postings on its own. A thread whose candidate does
      // not fit the buffer parks it and stops; the CTA meets at the barrier below, flushes once, and the loop resumes.
      // Driver lists are swept one after the other, so everything that depends on the leading list (what to probe,
      // where its postings live, who owns a doc) is loop invariant.
      int ct = 0;           // driver slot this thread is working on
      uint32_t cb = 0;      // first posting (of the slot's run segment) of the thread's next round
      uint64_t park[kR];
      uint32_t pmask = 0;   // parked candidates of this thread
      // generic queries: per-warp queue of (doc, tf word) pairs awaiting clause evaluation (lives in the bound table's
      // shared memory, which only pure disjunctions use); qn is warp-uniform
      uint2* const wq = reinterpret_cast<uint2*>(sm.ubt) + (tid >> 5) * 64;
      static_assert(kUbt * sizeof(float) >= (kThreads / 32) * 64 * sizeof(uint2), "warp queues do not fit the bound table");
      int qn = 0;
      const uint32_t q_req = sm.q.req_term_mask, q_not = sm.q.not_term_mask;
      auto drain = [&](int n_take) -> bool {   // evaluates the last n_take (<= 32) queued pairs; true: a lane had to park its hit
        __syncwarp();
        bool parked = false;
        const bool have = lane < n_take;
        const uint2 e = have ? wq[qn - n_take + lane] : make_uint2(0u, 0u);
        qn -= n_take;
        float score;
        if (have && evaluate_doc(L, sm, (int32_t)e.x, e.y, &score)) {
          const int32_t d = (int32_t)e.x;
          ++my_hits;
          if (L.aggs) agg_collect(*L.aggs, L.ix, qi, d);   // additional collectors see every matching doc
          uint64_t entry;
          if (L.sort_kind == NRTGPU_SORT_RELEVANCE) entry = make_key(score, d);
          else {   // TopFieldCollector: the key is the doc's sort value (order-preserving code), ties by doc id
            uint32_t code = 0;
            if (L.sort_kind == NRTGPU_SORT_COLUMN) { code = __ldg(L.sort_codes + d); if (code == 0u) code = sort_missing; }
            entry = ((uint64_t)sort_hi(L.sort_kind, L.sort_reverse, code, d) << 32) | (uint32_t)(~(uint32_t)d);
          }
          const unsigned long long th = sm.theta;
          if (entry > th && !(has_after && !(entry < after_key)) && !NRT_KNOCK(4)) {
            const int p = atomicAdd(&sm.cand_count, 1);
            if (p < kCand) sm.cand[p] = entry;
            else { park[0] = entry; pmask |= 1u; parked = true; }
          }
        }
        __syncwarp();
        return parked;
      };
      for (;;) {
        bool full = false;
#pragma unroll
        for (int j = 0; j < kR; ++j)
          if ((pmask >> j) & 1u) {
            const int p = atomicAdd(&sm.cand_count, 1);
            if (p < kCand) { sm.cand[p] = park[j]; pmask &= ~(1u << j); } else full = true;
          }
        if (!kSimple) {   // the warp moves together (its queue operations are collective); a queue left >= 32 by a full buffer is
          full = __any_sync(0xffffffffu, full);   // brought below 32 before the sweep pushes again (capacity 64)
          while (!full && qn >= 32) full = __any_sync(0xffffffffu, drain(32));
        }
        const int ct_end = NRT_KNOCK(8) ? 0 : (dense ? n_term + 1 : n_term);   // dense: one more "list" = every doc of the run
        while (!full && ct < ct_end) {
          const bool t_dense = ct == n_term;
          const uint32_t n_t = t_dense ? (uint32_t)(sm.run_d1 - sm.run_d0) : (((drv_mask >> ct) & 1u) ? sm.s_rb[ct] - sm.s_ra[ct] : 0u);
          if (cb >= n_t) { ++ct; cb = 0; continue; }
          const int t = t_dense ? 0 : ct;
          const uint32_t need = t_dense ? ((n_term >= 32) ? 0xffffffffu : ((1u << n_term) - 1u)) : sm.s_need[t];
          const uint32_t need_plane = NRT_KNOCK(1) ? 0u : need & plane_mask, need_long = NRT_KNOCK(2) ? 0u : need & long_mask,
                         need_short = NRT_KNOCK(2) ? 0u : need & short_mask, need_glob = need & global_mask;
          const bool t_staged = !t_dense && (((long_mask | short_mask) >> t) & 1u);
          const bool t_ess = (ess_mask >> t) & 1u;
          const uint32_t candbelow = t_dense ? 0u : sm.s_candbelow[t], cntbefore = t_dense ? 0u : sm.s_cntbefore[t];
          const int32_t dense_d0 = sm.run_d0;
          const int32_t* sdoc_t = sm.sdocs + ((int)sm.s_ra[t] + sm.s_sdelta[t]);   // posting x of the run segment: sdoc_t[x]
          const uint8_t* sf8_t = sm.sf8 + ((int)sm.s_ra[t] + sm.s_sdelta[t]);
          const int32_t* gdoc_t = sm.s_gdocs[t] + sm.s_ra[t];
          const uint8_t* gf8_t = sm.s_gf8[t] + sm.s_ra[t];
          const uint32_t tshift = t_dense ? 0u : 8u * (uint32_t)t;
          // software pipeline: the postings of the NEXT round are fetched before the current round is processed
          // (generic pointers: one load path for staged -- shared memory -- and plane / global -- HBM -- driver lists)
          const int32_t* dptr = t_staged ? sdoc_t : gdoc_t;
          const uint8_t* fptr = t_staged ? sf8_t : gf8_t;
          int32_t nd[kR]; uint32_t nf[kR];
#pragma unroll
          for (int j = 0; j < kR; ++j) {
            const uint32_t x = cb + (uint32_t)(j * kThreads + tid);
            nd[j] = -1; nf[j] = 0;   // doc -1: no posting
            if (x < n_t) {
              if (t_dense) nd[j] = dense_d0 + (int32_t)x; else { nd[j] = dptr[x]; nf[j] = fptr[x]; }
            }
          }
          while (!full && cb < n_t) {
            const unsigned long long theta = sm.theta;
            const float theta_s = theta ? key_score(theta) : -INFINITY;
            int32_t doc[kR]; uint32_t word[kR];
            uint32_t pbyte[kR][kT];
#pragma unroll
            for (int j = 0; j < kR; ++j) { doc[j] = nd[j]; word[j] = nf[j] << tshift; }
            // plane gathers of every posting of the round (2-bit tf codes, all in flight together)
            // (lanes without a posting gather byte 0 of the plane: harmless, and the branches stay CTA-uniform)
#pragma unroll
            for (int j = 0; j < kR; ++j) {
              const uint32_t d4 = (uint32_t)max(doc[j], 0) >> 2;
              pbyte[j][0] = 0u; pbyte[j][1] = 0u; pbyte[j][2] = 0u; pbyte[j][3] = 0u;
              if (need_plane & 1u) pbyte[j][0] = (uint32_t)__ldg(pl0 + d4);
              if (need_plane & 2u) pbyte[j][1] = (uint32_t)__ldg(pl1 + d4);
              if (need_plane & 4u) pbyte[j][2] = (uint32_t)__ldg(pl2 + d4);
              if (need_plane & 8u) pbyte[j][3] = (uint32_t)__ldg(pl3 + d4);
            }
            // next round's postings
            {
              const uint32_t nb = cb + (uint32_t)(kR * kThreads);
#pragma unroll
              for (int j = 0; j < kR; ++j) {
                const uint32_t x = nb + (uint32_t)(j * kThreads + tid);
                nd[j] = -1; nf[j] = 0;
                if (x < n_t) {
                  if (t_dense) nd[j] = dense_d0 + (int32_t)x; else { nd[j] = dptr[x]; nf[j] = fptr[x]; }
                }
              }
            }
            // searches of the staged lists (shared memory; overlaps the gathers)
            if (need_long | need_short | need_glob) {
#pragma unroll
              for (int j = 0; j < kR; ++j) {
                if (doc[j] < 0) continue;
                const int g = (doc[j] - slice_base) >> kLogGran;
#pragma unroll
                for (int u = 0; u < kT; ++u) {
                  uint32_t b = 0;
                  if ((need_long >> u) & 1u) {
                    const uint32_t lo = max(sm.gb[u][g], sm.s_ra[u]), hi = min(sm.gb[u][g + 1], sm.s_rb[u]);
                    if (hi > lo) b = probe_smem(sm, (int)lo + sm.s_sdelta[u], (int)hi + sm.s_sdelta[u], doc[j]);
                  } else if ((need_short >> u) & 1u) {
                    if (sm.s_rb[u] > sm.s_ra[u]) b = probe_smem(sm, (int)sm.s_ra[u] + sm.s_sdelta[u], (int)sm.s_rb[u] + sm.s_sdelta[u], doc[j]);
                  } else if ((need_glob >> u) & 1u) {
                    if (sm.s_rb[u] > sm.s_ra[u]) b = probe_global(sm.s_gdocs[u], sm.s_gf8[u], sm.s_ra[u], sm.s_rb[u], doc[j]);
                  }
                  word[j] |= b << (8 * u);
                }
              }
            }
            // ownership, hit count, bound test, append (pure disjunctions; the generic path follows)
#pragma unroll
            for (int j = 0; kSimple && j < kR; ++j) {
              if (doc[j] < 0) continue;
              uint32_t v = word[j];
              if (need_plane) {
                // the four gathered bytes side by side; the doc's 2-bit code of every plane with one shift and one mask
                // (bits shifted in from the neighbouring byte fall outside the mask); code 3 = "three or more" becomes
                // kTfInexact (resolved when the doc is scored)
                const uint32_t raw = pbyte[j][0] | (pbyte[j][1] << 8) | (pbyte[j][2] << 16) | (pbyte[j][3] << 24);
                const uint32_t codes = (raw >> (((uint32_t)doc[j] & 3u) * 2u)) & 0x03030303u;
                const uint32_t sat = __vcmpeq4(codes, 0x03030303u);   // 0xff where the code saturated
                v |= (codes & ~sat) | (sat & (kTfInexact * 0x01010101u));
              }
              uint64_t entry;
              if (kSimple) {
                if (live && !((live[doc[j] >> 5] >> (doc[j] & 31)) & 1u)) continue;   // deleted docs are neither counted nor collected
                if ((v & cntbefore) == 0u) ++my_hits;
                if (!t_ess || (v & candbelow) != 0u) continue;   // counted only / emitted by a lower list
                if (sm.ubt[__dp4a(__vminu4(v, 0x05050505u), 0xD8240601u, 0u)] < theta_s) continue;   // cannot reach the top-k
                entry = ((uint64_t)v << 32) | (uint32_t)doc[j];   // scored at the next flush
              } else {
                continue;   // (generic queries: the survivors of the round are queued below and evaluated a full warp at a time)
              }
              if (NRT_KNOCK(4)) continue;
              const int p = atomicAdd(&sm.cand_count, 1);
              if (p < kCand) sm.cand[p] = entry;
              else { park[j] = entry; pmask |= 1u << j; full = true; }
            }
            if (!kSimple) {
              // Generic queries: most driver postings fail the other required lists, so the clause evaluation (norm and
              // doc-value gathers, BM25) would run with a handful of lanes. The survivors of the cheap tests -- not owned by
              // a lower list, every required term present, no excluded term -- go to a per-warp queue and are evaluated
              // 32 at a time.
#pragma unroll
              for (int j = 0; j < kR; ++j) {
                uint32_t v = word[j];
                if (need_plane) {
                  const uint32_t raw = pbyte[j][0] | (pbyte[j][1] << 8) | (pbyte[j][2] << 16) | (pbyte[j][3] << 24);
                  const uint32_t codes = (raw >> (((uint32_t)max(doc[j], 0) & 3u) * 2u)) & 0x03030303u;
                  const uint32_t sat = __vcmpeq4(codes, 0x03030303u);
                  v |= (codes & ~sat) | (sat & (kTfInexact * 0x01010101u));
                }
                const uint32_t pres = v2::presence4(v);
                const bool surv = doc[j] >= 0 && (v & candbelow) == 0u && (pres & q_req) == q_req && (pres & q_not) == 0u;
                const unsigned bal = __ballot_sync(0xffffffffu, surv);
                if (surv) wq[qn + __popc(bal & ((1u << lane) - 1u))] = make_uint2((uint32_t)doc[j], v);
                qn += __popc(bal);
                while (!full && qn >= 32) full = __any_sync(0xffffffffu, drain(32));   // (a parked hit stops the warp: park[0] is free whenever drain runs)
              }
            }
            cb += kR * kThreads;
            if (kStats && tid == 0) ++dbg_rounds;
          }
        }
        if (!kSimple) {   // the rest of the warp's queue (a partial warp), unless the buffer is full: then after the flush
          while (!full && qn > 0) full = __any_sync(0xffffffffu, drain(min(qn, 32)));
        }
        __syncthreads();
        if (sm.cand_count <= kCand) break;   // nobody is parked (the count passes kCand only through a failed append)
        const long long tf = kStats ? clock64() : 0ll;
        flush_candidates<kSimple>(L, sm, norms0, n_term, has_after, after_key, L.top_k, &L.theta[qi]);
        if (kStats) dbg_tflush += clock64() - tf;
        if (kStats) ++dbg_flush;
      }
      g0 = g1;
      first_run = false;
      __syncthreads();   // R0: every thread is done with the stage before the next run overwrites it
    }

    // ---------------- finish the work item (the slice merge sorts, so only a full buffer needs ordering here)
    __syncthreads();
    if (kSimple ? sm.cand_count > sm.n_keys : sm.cand_count > L.top_k) {
      const long long tf = kStats ? clock64() : 0ll;
      flush_candidates<kSimple>(L, sm, norms0, n_term, has_after, after_key, L.top_k, &L.theta[qi]);
      if (kStats) dbg_tflush += clock64() - tf;
      if (kStats) ++dbg_flush;
    }
    const int keep = sweep_warm ? 0 : min(sm.cand_count, L.top_k);
    const int out_list = (wflags & 5) ? L.n_lists - 1 : slice * L.parts_max + part * kfine;
    uint64_t* out = L.slice_keys + ((size_t)qi * L.n_lists + out_list) * L.top_k;
    for (int i = tid; i < keep; i += kThreads) out[i] = sm.cand[i];
    if (tid == 0) L.slice_cnt[(size_t)qi * L.n_lists + out_list] = keep;
    for (int o = 16; o > 0; o >>= 1) my_hits += __shfl_xor_sync(0xffffffffu, my_hits, o);
    if (lane == 0 && my_hits && !sweep_warm) atomicAdd(&L.total_hits[qi], (unsigned long long)my_hits);
    if (kStats && tid == 0) {
      atomicAdd(&L.stats[0], 1ull);
      atomicAdd(&L.stats[1], (unsigned long long)(clock64() - t_start));
      atomicAdd(&L.stats[2], (unsigned long long)dbg_runs);
      atomicAdd(&L.stats[3], dbg_post);
      atomicAdd(&L.stats[4], (unsigned long long)dbg_flush);
      atomicAdd(&L.stats[5], (unsigned long long)dbg_staged);
      atomicAdd(&L.stats[6], (unsigned long long)(t_setup - t_start));
      atomicAdd(&L.stats[7], (unsigned long long)dbg_rounds);
      const unsigned long long cyc = (unsigned long long)(clock64() - t_start);
      atomicMax(&L.stats[8], cyc);
      if (wflags & 5) { atomicAdd(&L.stats[11], 1ull); atomicAdd(&L.stats[12], cyc); }
      atomicAdd(&L.stats[13], (unsigned long long)dbg_tflush); atomicAdd(&L.stats[14], (unsigned long long)dbg_twait);
    }
  }
  if (kStats && tid == 0) {
    const unsigned long long busy = (unsigned long long)(clock64() - t_cta);
    atomicAdd(&L.stats[9], busy); atomicMax(&L.stats[10], busy);
  }
}

// postings of every (query, term slot) below each part boundary of every slice (parts_max equal granule ranges per
// slice), below the end of the shard, and below the warm-up boundary of slice 0
struct SliceBoundsLaunch {
  DevIndexView ix;
  const DevClause* clauses;
  const DevQuery* queries;
  int32_t nq, n_slices, slice_gran, n_gran, parts_max;
  uint32_t* sbounds;   // [nq][kT][n_slices * parts_max + 2]
};

__global__ void slice_bounds_kernel(SliceBoundsLaunch B) {
  const int n_b = B.n_slices * B.parts_max;
  const int per_slot = n_b + 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B.nq * kT * per_slot) return;
  const int q = (int)(i / (kT * per_slot)), s = (int)((i / per_slot) % kT), e = (int)(i % per_slot);
  const DevQuery dq = B.queries[q];
  const int fine = (B.slice_gran + B.parts_max - 1) / B.parts_max;
  int64_t gran;
  if (e < n_b) gran = (int64_t)(e / B.parts_max) * B.slice_gran + min(B.slice_gran, (e % B.parts_max) * fine);
  else if (e == n_b) gran = B.n_gran;
  else gran = min(kWarmGran, B.slice_gran);
  if (gran > B.n_gran) gran = B.n_gran;
  uint32_t out = 0;
  for (int c = 0; c < dq.n_clauses; ++c) {
    const DevClause cl = B.clauses[dq.clause_begin + c];
    if (cl.kind != NRTGPU_TERM || cl.slot != s) continue;
    if (cl.gran_row >= 0) { out = __ldg(B.ix.gran_tab + (size_t)cl.gran_row * (size_t)(B.n_gran + 1) + gran); break; }
    const int64_t target64 = gran << kLogGran;
    const int32_t target = target64 > (int64_t)B.ix.n_docs ? B.ix.n_docs : (int32_t)target64;
    const int32_t* docs = B.ix.post_docs + cl.post_base;
    int lo = 0, hi = cl.n_post;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(docs + mid) < target) lo = mid + 1; else hi = mid; }
    out = (uint32_t)lo;
    break;
  }
  B.sbounds[i] = out;
}

}  // namespace v3
}  // namespace nrtgpu
