// nrtgpu.cu -- C ABI (include/nrtgpu.h) of the B200 query-execution engine: context, HBM index image,
// batch compilation, kernel launches. No CPU fallback: every entry point needs a CUDA device.
#include "../../include/nrtgpu.h"
#include "bool_kernel.cuh"
#include "stream_kernel.cuh"
#include "probe_kernel.cuh"
#include "sort_kernel.cuh"
#include "collect_kernel.cuh"
#include "knn_kernel.cuh"
#include "hybrid_kernel.cuh"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <thread>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

namespace nrtgpu {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace nrtgpu
using namespace nrtgpu;

#define NRT_FAIL(code, msg) do { set_error(msg); return (code); } while (0)


// index-time impacts: max over a term's postings of x = tf * cache[norm] (what Lucene keeps as competitive (freq, norm)
// pairs in its skip data); the BM25 score is monotone in x, so score(weight, max x) bounds the whole list.
// One warp per term (lists of thousands of postings take a whole CTA's worth of iterations, the millions of tiny lists one
// each): no per-posting dictionary search, no atomics.
__global__ void term_max_x_kernel(const int64_t* __restrict__ term_off, int n_terms, const int32_t* __restrict__ term_field,
                                  const int32_t* __restrict__ docs, const uint8_t* __restrict__ f8,
                                  const int64_t* __restrict__ exc_pos, const int32_t* __restrict__ exc_freq, int n_exc,
                                  const uint8_t* const* __restrict__ norms, const float* __restrict__ caches,
                                  float* __restrict__ out) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_terms) return;
  const int t = (int)warp, f = term_field[t];
  const uint8_t* nrm = norms[f];
  const float* cache = caches + f * 256;
  float m = 0.0f;
  for (int64_t p = term_off[t] + lane; p < term_off[t + 1]; p += 32) {
    float freq = (float)f8[p];
    if (f8[p] == 255) {
      int a = 0, b = n_exc;
      while (a < b) { const int mid = (a + b) >> 1; if (exc_pos[mid] < p) a = mid + 1; else b = mid; }
      if (a < n_exc && exc_pos[a] == p) freq = (float)exc_freq[a];
    }
    m = fmaxf(m, __fmul_rn(freq, cache[nrm ? nrm[docs[p]] : 1]));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) out[t] = m;
}

// 2-bit planes: four docs per byte, min(tf, 3) each
__global__ void plane_pack2_kernel(const uint8_t* __restrict__ planes, int64_t n_bytes_out, uint8_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bytes_out) return;
  const uchar4 b = reinterpret_cast<const uchar4*>(planes)[i];
  out[i] = (uint8_t)(min((int)b.x, 3) | (min((int)b.y, 3) << 2) | (min((int)b.z, 3) << 4) | (min((int)b.w, 3) << 6));
}

// dense tf plane of one term: plane[doc] = min(freq, 255) for every posting of the term (the plane is zeroed first)
__global__ void plane_fill_kernel(const int32_t* __restrict__ docs, const uint8_t* __restrict__ f8, int64_t n,
                                  uint8_t* __restrict__ plane) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) plane[docs[i]] = f8[i];
}

namespace {

// ---- SmallFloat.byte4ToInt (Lucene) : norm byte -> field length, for the BM25 length table ----
int64_t int4_to_long(int i) {
  int64_t bits = i & 0x07;
  int shift = (i >> 3) - 1;
  return shift == -1 ? bits : ((bits | 0x08) << shift);
}
int32_t byte4_to_int(uint8_t b) {
  const int kFree = 24;  // 255 - longToInt4(Integer.MAX_VALUE)
  return b < kFree ? (int32_t)b : (int32_t)(kFree + int4_to_long((int)b - kFree));
}
// BM25Similarity.scorer(): cache[i] = 1f / (k1 * ((1 - b) + b * LENGTH_TABLE[i] / avgdl)), float ops
void bm25_cache(float k1, float b, float avgdl, float* cache) {
  for (int i = 0; i < 256; ++i) {
    volatile float t = b * (float)byte4_to_int((uint8_t)i);
    t = t / avgdl;
    t = (1.0f - b) + t;
    t = k1 * t;
    cache[i] = 1.0f / t;
  }
}
float bm25_idf(int64_t df, int64_t doc_count) {
  return (float)std::log(1.0 + ((double)doc_count - (double)df + 0.5) / ((double)df + 0.5));
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0, cap = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t count) {   // keeps the allocation when it is already large enough (workspace reuse)
    n = count;
    if (count <= cap) return NRTGPU_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    NRT_CUDA_TRY(cudaMalloc((void**)&p, count * sizeof(T)));
    cap = count;
    return NRTGPU_OK;
  }
  int upload_async(const T* h, size_t count, cudaStream_t st) {
    int rc = alloc(count);
    if (rc) return rc;
    if (count) NRT_CUDA_TRY(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st));
    return NRTGPU_OK;
  }
  int upload(const T* h, size_t count) {
    int rc = alloc(count);
    if (rc) return rc;
    if (count) NRT_CUDA_TRY(cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice));
    return NRTGPU_OK;
  }
  size_t bytes() const { return n * sizeof(T); }
};

}  // namespace

struct nrtgpu_ctx {
  int device = 0;
  int sm_count = 0;
  bool engine_stream = false;   // NRTGPU_ENGINE=stream: round-1 window/stream kernel for every <= 4-term query (A/B runs)
  std::mutex hyb_mu;             // O(k) hybrid stages share one pooled device scratch (no cudaMalloc per call)
  DevBuf<int32_t> hyb_scratch;
  int64_t item_postings = 32768; // NRTGPU_ITEM_POSTINGS: floor of the postings a (query, slice) may hold before it is split into 2..16 parts
  int64_t item_share_full = 32;  // NRTGPU_ITEM_SHARE_FULL: the same for launches that visit every posting (ScoreMode.COMPLETE, generic clause evaluation)
  int64_t item_share = 12;       // NRTGPU_ITEM_SHARE: ... and it is split when it exceeds 1/share of the postings per resident CTA
  int64_t warm_min_docs = 8ll * v2::kWarmGran * v2::kGran;   // NRTGPU_WARM_MIN_DOCS: shards below this size run without warm-up items
  bool warm_sweep = true;        // NRTGPU_WARM=docs: warm-up items sweep the first 32K docs (round-2 first style) instead of the rarest list
  int slice_gran = 512;          // NRTGPU_SLICE_GRAN: granules (1024 docs) per slice of the probe kernel, <= v3::kMaxSliceGran
  int probe_cfg = 0;             // NRTGPU_PROBE_CFG: 0 auto, 1 always A (3 CTAs / SM), 2 always B (4 CTAs / SM)
  bool order_lpt = false;        // NRTGPU_ORDER=lpt: query-major work order, longest query first
  bool order_by_cost = false;   // NRTGPU_ORDER=cost: round-1 work order (longest query first) instead of plane clusters
  bool debug_modes = false;     // NRTGPU_DEBUG_MODES=1: per-launch kernel statistics on stderr (adds a stream synchronisation)
};

// which launch configuration of the probe kernel a launch takes (see probe_kernel.cuh kCtasA / kCtasB)
static inline bool ix_ctx_probe_cfg(const nrtgpu_ctx* c, bool visits_everything) { return c->probe_cfg == 2 || (c->probe_cfg == 0 && visits_everything); }

struct nrtgpu_index {
  nrtgpu_ctx* ctx = nullptr;
  int32_t n_docs = 0, doc_base = 0, n_terms = 0, n_fields = 0, n_columns = 0;
  // host-side dictionary
  std::vector<int64_t> term_off;
  std::vector<int32_t> term_field;
  std::vector<int64_t> term_df;
  std::vector<float> term_max_x;
  std::vector<int32_t> term_plane;   // dense tf plane per term, -1 for all but the densest terms
  std::vector<int32_t> term_gran;    // row of the granule offset table per term, -1 for short lists
  std::vector<int64_t> field_doc_count, field_sum_ttf;
  std::vector<uint8_t> field_has_norms;
  std::vector<float> field_k1, field_b;
  // device image
  DevBuf<int32_t> post_docs;
  DevBuf<uint8_t> post_f8;
  DevBuf<int64_t> exc_pos;
  DevBuf<int32_t> exc_freq;
  std::vector<std::unique_ptr<DevBuf<uint8_t>>> norms;
  DevBuf<const uint8_t*> norms_ptrs;
  DevBuf<float> caches;
  DevBuf<uint32_t> gran_tab;        // [n_rows][n_gran + 1] index-time granule offsets (skip data) of the long lists
  int32_t gran_n = 0;
  DevBuf<uint8_t> dense_tf;         // [n_planes][dense_stride]: direct-address tf bytes of the densest terms
  DevBuf<uint8_t> dense_tf2;        // [n_planes][dense_stride / 4]: the same planes at 2 bits per doc (min(tf, 3))
  int64_t dense_stride = 0;
  int32_t n_planes = 0;
  DevBuf<uint8_t> field_min_norm;   // smallest non-zero norm byte per field (0 byte = doc lacks the field)
  std::vector<std::unique_ptr<DevBuf<int64_t>>> col64;
  std::vector<std::unique_ptr<DevBuf<int32_t>>> col32;
  std::vector<std::unique_ptr<DevBuf<uint8_t>>> col_has;
  std::vector<std::unique_ptr<DevBuf<uint32_t>>> col_code;      // per column: order-preserving sort code per doc (sort_kernel.cuh)
  std::vector<std::unique_ptr<DevBuf<uint64_t>>> col_distinct;  // per column: sorted distinct values (sortable domain)
  std::vector<int32_t> col_n_distinct;
  DevBuf<const int64_t*> col64_ptrs;
  DevBuf<const int32_t*> col32_ptrs;
  DevBuf<const uint8_t*> col_has_ptrs;
  std::vector<std::unique_ptr<DevBuf<int64_t>>> colmv_off;   // multi-valued columns: per-doc offsets (values sit in col64)
  DevBuf<const int64_t*> colmv_off_ptrs, colmv_val_ptrs;
  std::vector<uint8_t> col_multi;                            // [n_columns] 1 = multi-valued
  DevBuf<uint32_t> live_bits;
  // vectors
  int32_t vec_dims = 0, vec_sim = 0, vec_count = 0;
  bool vec_is_byte = false;   // byte vector field (ByteVectorFieldDef): same image, byte score mapping
  DevBuf<float> vectors;
  DevBuf<__nv_bfloat16> vec_bf16;   // bf16 copy of the corpus for the tensor-core candidate stage (dims % 8 == 0)
  CUtensorMap vec_tmap;             // TMA tensor map over vec_bf16 (256-row boxes)
  CUtensorMap vec_tmap128;          // ... (128-row boxes: the double-buffered GEMM's corpus tile)
  DevBuf<float2> vec_ab;            // per-vector (a, b) of the approximate score a * dot + b
  bool vec_tc = false;
  float vec_dmax = 0.0f;    // largest vector magnitude (error bound of the kNN candidate-stage certificate)
  int32_t knn_last_uncertified = 0;   // queries of the last kNN call that took the exact fallback
  DevBuf<float> vec_norm2;  // per-vector squared magnitude (double-accumulated, stored float) for cosine
  DevBuf<int32_t> vec_docs;
  int64_t device_bytes = 0;
  // reusable batch workspaces of the one-shot entry point (nrtgpu_search_bool), one per concurrent caller
  std::mutex ws_mu;
  std::mutex knn_mu;          // one kNN call at a time per index: they share knn_scratch
  std::mutex fetch_mu;        // fetch-phase scratch
  DevBuf<int32_t> f_cols, f_docs; DevBuf<int64_t> f_vals; DevBuf<uint8_t> f_has;
  KnnScratch knn_scratch;
  std::vector<nrtgpu_batch*> ws_free;
  ~nrtgpu_index();

  DevIndexView view() const {
    DevIndexView v;
    v.n_docs = n_docs; v.doc_base = doc_base;
    v.post_docs = post_docs.p; v.post_f8 = post_f8.p;
    v.exc_pos = exc_pos.p; v.exc_freq = exc_freq.p; v.n_exc = (int32_t)exc_pos.n;
    v.norms = norms_ptrs.p; v.caches = caches.p;
    v.col64 = col64_ptrs.p; v.col32 = col32_ptrs.p; v.col_has = col_has_ptrs.p;
    v.colmv_off = colmv_off_ptrs.p; v.colmv_val = colmv_val_ptrs.p;
    v.live_bits = live_bits.p;
    v.dense_tf = dense_tf.p; v.dense_stride = dense_stride; v.dense_tf2 = dense_tf2.p;
    v.gran_tab = gran_tab.p; v.n_gran = gran_n;
    return v;
  }
};

struct nrtgpu_batch;
static void free_batch(nrtgpu_batch* b);
struct nrtgpu_batch {
  nrtgpu_index* ix = nullptr;
  int32_t nq = 0, top_k = 0, n_slices = 0, n_work = 0;
  int32_t n_lists = 0;         // per-query candidate lists the kernels fill: n_slices (+1: warm-up items of the stream path)
  int32_t n_work_simple = 0;   // the first n_work_simple work items belong to pure single-field term disjunctions
  // work list layout (<= 4-term batches): [probe simple | probe generic | stream (window kernel: no posting list can lead)]
  int32_t n_probe_simple = 0, n_probe_generic = 0, n_stream = 0;
  bool use_probe = false;
  DevBuf<uint32_t> sbounds;            // probe kernel: [nq][4][n_slices * parts_max + 2] part-boundary posting offsets
  DevBuf<unsigned int> work_counter;   // probe kernel: queue heads [2]
  DevBuf<unsigned long long> probe_stats;
  bool wide_slots = false;
  bool exhaustive = true;
  int64_t alg_postings = 0;
  DevBuf<DevClause> clauses;
  DevBuf<DevQuery> queries;
  DevBuf<int32_t> work_query, work_slice;
  DevBuf<uint32_t> gbounds;  // stream kernel: [nq][4][n_gran+1]
  DevBuf<float> qtables;     // stream kernel: [nq][kQTabFloats] score + bound tables
  DevBuf<unsigned long long> mode_stats;   // NRTGPU_DEBUG_MODES=1: cycles / work items per kernel mode
  DevBuf<int32_t> pruned;    // [nq] relation GTE flags
  DevBuf<int32_t> terminated; // [nq] terminateAfter cut the query short
  DevBuf<int32_t> timed_out;  // [nq] a work item of the query was skipped because the deadline had passed
  DevBuf<unsigned long long> clock0;  // [1] %globaltimer when the first work item of the run started
  std::vector<int32_t> h_flags;
  // sort-by-field (TopFieldCollector)
  int32_t sort_kind = 0, sort_column = 0, sort_reverse = 0;
  int64_t sort_missing_value = 0;
  DevBuf<int64_t> after_values; DevBuf<int32_t> after_docs; DevBuf<uint32_t> sort_missing_code;
  DevBuf<int64_t> out_sort_values;
  std::vector<int32_t> h_after_docs;
  // additional collectors (aggregations)
  std::vector<nrtgpu_aggregation> aggs;
  DevBuf<unsigned int> agg_counts[kMaxAggs];
  DevBuf<unsigned long long> agg_dvals[kMaxAggs];
  DevBuf<AggLaunch> agg_launch;
  DevBuf<int64_t> agg_keys; DevBuf<int32_t> agg_cnts, agg_n, agg_tot; DevBuf<long long> agg_other;
  // second pass of QueryRescorer (nrtgpu_score_docs / nrtgpu_rescore_query)
  DevBuf<int32_t> sd_docs, sd_counts; DevBuf<uint8_t> sd_match; DevBuf<float> sd_scores, sd_first;
  bool limits_active = false, disallow_partial = false;
  double timeout_sec = 0.0;
  long long deadline_ns = 0;       // budget from the first work item on (0: none)
  int64_t ta_scalar = 0;           // terminateAfter (0: none)
  int64_t terminate_after_max_recall = 0;
  std::vector<DevClause> h_dc; std::vector<DevQuery> h_dq; std::vector<int32_t> h_wq, h_ws;   // host copies the async uploads read
  int32_t slice_docs = 0;
  DevBuf<unsigned long long> known_hits;        // probe kernel: docs known to match per query (0: unknown)
  std::vector<unsigned long long> h_known;
  int32_t parts_max = 1;       // probe kernel: parts a (query, slice) work item may be split into (power of two)
  int64_t threshold = INT32_MAX;
  int32_t n_gran = 0;
  DevBuf<uint64_t> theta;
  DevBuf<unsigned long long> total_hits;
  DevBuf<uint64_t> slice_keys;
  DevBuf<int32_t> slice_cnt;
  DevBuf<int32_t> out_docs;
  DevBuf<float> out_scores;
  DevBuf<int32_t> out_counts;
  static constexpr int kEvRing = 64;
  cudaEvent_t ev[kEvRing][3] = {};
  int runs_recorded = 0;   // since the last timing reset
  bool ran = false;
  // optional caller-provided device output buffers (e.g. torch tensors feeding the NCCL all-gather)
  int32_t* bound_docs = nullptr; float* bound_scores = nullptr; int32_t* bound_counts = nullptr;
  long long* bound_total = nullptr; int32_t* bound_flags = nullptr;   // packed record (nrtgpu_batch_bind_packed)
  int32_t* o_docs() { return bound_docs ? bound_docs : out_docs.p; }
  float* o_scores() { return bound_scores ? bound_scores : out_scores.p; }
  int32_t* o_counts() { return bound_counts ? bound_counts : out_counts.p; }
  ~nrtgpu_batch() { for (auto& r : ev) for (auto& e : r) if (e) cudaEventDestroy(e); }
};

static void free_batch(nrtgpu_batch* b) { delete b; }

// index-time impacts: max over a term's postings of tf * cache[norm]; depends on the index-wide avgdl through cache[]
static int compute_term_max_x(nrtgpu_index* ix) {
  ix->term_max_x.assign((size_t)ix->n_terms, 0.0f);
  const int64_t P = ix->n_terms ? ix->term_off[(size_t)ix->n_terms] : 0;
  if (P <= 0) return NRTGPU_OK;
  int rc;
  DevBuf<int64_t> d_off; DevBuf<int32_t> d_tf; DevBuf<float> d_mx;
  if ((rc = d_off.upload(ix->term_off.data(), ix->term_off.size()))) return rc;
  if ((rc = d_tf.upload(ix->term_field.data(), ix->term_field.size()))) return rc;
  if ((rc = d_mx.alloc((size_t)ix->n_terms))) return rc;
  const int64_t threads = (int64_t)ix->n_terms * 32;
  term_max_x_kernel<<<(unsigned)((threads + 255) / 256), 256>>>(d_off.p, ix->n_terms, d_tf.p, ix->post_docs.p, ix->post_f8.p,
                                                                 ix->exc_pos.p, ix->exc_freq.p, (int)ix->exc_pos.n, ix->norms_ptrs.p,
                                                                 ix->caches.p, d_mx.p);
  NRT_CUDA_TRY(cudaGetLastError());
  NRT_CUDA_TRY(cudaMemcpy(ix->term_max_x.data(), d_mx.p, (size_t)ix->n_terms * sizeof(float), cudaMemcpyDeviceToHost));
  return NRTGPU_OK;
}

static int upload_live_docs(nrtgpu_index* ix, const uint8_t* live_docs) {
  if (!live_docs) { ix->live_bits.n = 0; if (ix->live_bits.p) { cudaFree(ix->live_bits.p); ix->live_bits.p = nullptr; ix->live_bits.cap = 0; } return NRTGPU_OK; }
  std::vector<uint32_t> bits(((size_t)ix->n_docs + 31) / 32, 0u);
  for (int32_t i = 0; i < ix->n_docs; ++i) if (live_docs[i]) bits[(size_t)i >> 5] |= 1u << (i & 31);
  return ix->live_bits.upload(bits.data(), bits.size());
}
extern "C" {
static int batch_fetch_aggs(nrtgpu_batch* b, cudaStream_t st, const nrtgpu_aggregation_result* out);
static int batch_set_limits(nrtgpu_batch* b, const nrtgpu_search_limits* lim, cudaStream_t st);
static int batch_fetch_impl(nrtgpu_batch* b, void* stream_, int32_t* out_docs, float* out_scores, int32_t* out_counts,
                            int64_t* out_total_hits, uint8_t* out_relation, uint8_t* out_hit_timeout, uint8_t* out_terminated_early);
}
nrtgpu_index::~nrtgpu_index() { for (auto* b : ws_free) free_batch(b); }

extern "C" {

const char* nrtgpu_last_error(void) { return g_last_error.c_str(); }
int nrtgpu_version(void) { return 1; }

int nrtgpu_init(int device_id, nrtgpu_ctx** out) {
  if (!out) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_init: out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    NRT_FAIL(NRTGPU_ERR_CUDA, std::string("nrtgpu_init: no CUDA device (") + cudaGetErrorString(e) +
                                  "); this engine has no CPU fallback");
  if (device_id < 0 || device_id >= n) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_init: bad device id");
  NRT_CUDA_TRY(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  NRT_CUDA_TRY(cudaGetDeviceProperties(&prop, device_id));
  if (prop.major < 10) NRT_FAIL(NRTGPU_ERR_CUDA, "nrtgpu_init: device is not sm_100 class (kernels are built for sm_100a only)");
  std::unique_ptr<nrtgpu_ctx> c(new nrtgpu_ctx);   // released to the caller only when every attribute call succeeded
  c->device = device_id;
  c->sm_count = prop.multiProcessorCount;
  { const char* e = getenv("NRTGPU_ENGINE"); c->engine_stream = e && std::strcmp(e, "stream") == 0; }
  c->debug_modes = getenv("NRTGPU_DEBUG_MODES") != nullptr;
  { const char* e = getenv("NRTGPU_PROBE_CFG"); c->probe_cfg = e ? atoi(e) : 0; }
  { const char* e = getenv("NRTGPU_WARM_MIN_DOCS"); if (e && atoll(e) > 0) c->warm_min_docs = atoll(e); }
  { const char* e = getenv("NRTGPU_WARM"); c->warm_sweep = !(e && std::strcmp(e, "docs") == 0); }
  { const char* e = getenv("NRTGPU_SLICE_GRAN"); if (e && atoi(e) >= 64) c->slice_gran = std::min(atoi(e), (int)v3::kMaxSliceGran); }
  { const char* e = getenv("NRTGPU_ITEM_POSTINGS"); if (e && atoll(e) > 0) c->item_postings = atoll(e); }
  { const char* e = getenv("NRTGPU_ITEM_SHARE"); if (e && atoll(e) > 0) c->item_share = atoll(e); }
  { const char* e = getenv("NRTGPU_ITEM_SHARE_FULL"); if (e && atoll(e) > 0) c->item_share_full = atoll(e); }
  { const char* e = getenv("NRTGPU_ORDER"); c->order_by_cost = e && (std::strcmp(e, "cost") == 0 || std::strcmp(e, "lpt") == 0); c->order_lpt = e && std::strcmp(e, "lpt") == 0; }
  NRT_CUDA_TRY(cudaFuncSetAttribute(bool_window_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(BoolSmem<uint32_t>)));
  NRT_CUDA_TRY(cudaFuncSetAttribute(bool_window_kernel<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(BoolSmem<uint64_t>)));
  NRT_CUDA_TRY(cudaFuncSetAttribute(v2::posting_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(v2::StreamSmem)));
  NRT_CUDA_TRY(cudaFuncSetAttribute(v2::posting_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(v2::StreamSmem)));
#define NRT_PROBE_ATTR(S, D) \
  NRT_CUDA_TRY(cudaFuncSetAttribute(v3::posting_probe_kernel<S, D, v3::kCtasA, v3::kStageA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(v3::ProbeSmemT<v3::kStageA>))); \
  NRT_CUDA_TRY(cudaFuncSetAttribute(v3::posting_probe_kernel<S, D, v3::kCtasB, v3::kStageB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(v3::ProbeSmemT<v3::kStageB>)));
  NRT_PROBE_ATTR(true, false) NRT_PROBE_ATTR(false, false) NRT_PROBE_ATTR(true, true) NRT_PROBE_ATTR(false, true)
#undef NRT_PROBE_ATTR
  NRT_CUDA_TRY(cudaFuncSetAttribute(tc::knn_gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kGemmSmem));
  NRT_CUDA_TRY(cudaFuncSetAttribute(tc::knn_gemm_bf16_db_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kGemm3Smem));
  NRT_CUDA_TRY(cudaFuncSetAttribute(tc::knn_gemm_bf16_256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kGemm2Smem));
  NRT_CUDA_TRY(cudaFuncSetAttribute(tc::knn_gemm_bf16_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kPGemmSmem));
  *out = c.release();
  return NRTGPU_OK;
}

void nrtgpu_shutdown(nrtgpu_ctx* ctx) { delete ctx; }

int nrtgpu_index_build(nrtgpu_ctx* ctx, const nrtgpu_shard_desc* d, nrtgpu_index** out) {
  if (!ctx || !d || !out) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: NULL argument");
  if (d->n_docs < 0 || d->n_terms < 0 || d->n_fields < 0 || d->n_columns < 0)
    NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: negative size");
  if (d->n_terms > 0 && (!d->term_off || d->n_fields < 1 || !d->field_doc_count || !d->field_sum_ttf))
    NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: terms need term_off and field statistics");
  NRT_CUDA_TRY(cudaSetDevice(ctx->device));
  std::unique_ptr<nrtgpu_index> ix(new nrtgpu_index);
  ix->ctx = ctx;
  ix->n_docs = d->n_docs; ix->doc_base = d->doc_base; ix->n_terms = d->n_terms;
  ix->n_fields = d->n_fields; ix->n_columns = d->n_columns;
  int rc;
  const int64_t P = d->n_terms ? d->term_off[d->n_terms] : 0;
  ix->term_off.assign(d->term_off, d->term_off + (d->n_terms ? d->n_terms + 1 : 0));
  ix->term_field.resize(d->n_terms);
  ix->term_df.resize(d->n_terms);
  for (int t = 0; t < d->n_terms; ++t) {
    int f = d->term_field ? d->term_field[t] : 0;
    if (f < 0 || f >= d->n_fields) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: term_field out of range");
    int64_t len = d->term_off[t + 1] - d->term_off[t];
    if (len < 0 || len > INT32_MAX) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: bad term_off");
    ix->term_field[t] = f;
    ix->term_df[t] = d->term_df ? d->term_df[t] : len;
  }
  if (d->n_terms > 0 && d->term_off[0] != 0) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: term_off[0] must be 0");
  if (P > 0 && (!d->post_docs || !d->post_freqs)) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: NULL postings");
  // every list strictly ascending and inside [0, n_docs): the kernels binary-search the lists and index per-doc arrays with them
  for (int t = 0; t < d->n_terms; ++t) {
    int32_t prev = -1;
    for (int64_t p = d->term_off[t]; p < d->term_off[t + 1]; ++p) {
      const int32_t doc = d->post_docs[p];
      if (doc <= prev || doc >= d->n_docs) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: post_docs must be strictly ascending per term and < n_docs");
      prev = doc;
    }
  }
  if (d->vec_dims < 0 || d->vec_count < 0) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: negative vector size");
  if (d->vec_dims > 0 && d->vec_count > 0) {
    if (!d->vec_docs && d->vec_count > d->n_docs) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: vec_count > n_docs with an identity ord -> doc map");
    if (d->vec_docs) for (int32_t i = 0; i < d->vec_count; ++i)
      if (d->vec_docs[i] < 0 || d->vec_docs[i] >= d->n_docs) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: vec_docs out of range");
  }
  ix->field_doc_count.assign(d->field_doc_count, d->field_doc_count + d->n_fields);
  ix->field_sum_ttf.assign(d->field_sum_ttf, d->field_sum_ttf + d->n_fields);
  // postings
  const size_t pad = 2 * (size_t)v2::kCH;   // whole TMA chunks may extend past the last posting
  if ((rc = ix->post_docs.alloc((size_t)P + pad))) return rc;
  NRT_CUDA_TRY(cudaMemset(ix->post_docs.p, 0x7f, ((size_t)P + pad) * sizeof(int32_t)));
  if (P) NRT_CUDA_TRY(cudaMemcpy(ix->post_docs.p, d->post_docs, (size_t)P * sizeof(int32_t), cudaMemcpyHostToDevice));
  {
    std::vector<uint8_t> f8((size_t)P + pad, 0);
    std::vector<int64_t> epos; std::vector<int32_t> efreq;
    for (int64_t p = 0; p < P; ++p) {
      int32_t f = d->post_freqs[p];
      if (f < 1) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: term frequency < 1");
      if (f >= 255) { f8[(size_t)p] = 255; epos.push_back(p); efreq.push_back(f); } else f8[(size_t)p] = (uint8_t)f;
    }
    if ((rc = ix->post_f8.upload(f8.data(), (size_t)P + pad))) return rc;
    if ((rc = ix->exc_pos.upload(epos.data(), epos.size()))) return rc;
    if ((rc = ix->exc_freq.upload(efreq.data(), efreq.size()))) return rc;
  }
  // dense tf planes: a term present in >= 1/64 of the docs also gets a direct-address byte per doc (like the bit-set
  // blocks Lucene's postings format keeps for dense blocks). A list that only needs LOOKUPS in a window (a
  // non-essential MAXSCORE list) is then one TMA copy of the window's bytes instead of a scatter of its postings.
  ix->term_plane.assign((size_t)d->n_terms, -1);
  if (d->n_docs >= 4096) {
    std::vector<int32_t> dense_terms;
    for (int32_t t = 0; t < d->n_terms; ++t)
      if ((d->term_off[t + 1] - d->term_off[t]) * 64 >= (int64_t)d->n_docs) dense_terms.push_back(t);
    const int64_t stride = (((int64_t)d->n_docs + 15) / 16) * 16 + 16;
    const size_t max_planes = std::min<size_t>(1024, (size_t)((8ll << 30) / stride));
    if (dense_terms.size() > max_planes) {   // keep the densest
      std::sort(dense_terms.begin(), dense_terms.end(), [&](int32_t a, int32_t b) {
        return d->term_off[a + 1] - d->term_off[a] > d->term_off[b + 1] - d->term_off[b]; });
      dense_terms.resize(max_planes);
    }
    if (!dense_terms.empty()) {
      ix->dense_stride = stride; ix->n_planes = (int32_t)dense_terms.size();
      if ((rc = ix->dense_tf.alloc((size_t)stride * dense_terms.size()))) return rc;
      NRT_CUDA_TRY(cudaMemset(ix->dense_tf.p, 0, ix->dense_tf.bytes()));
      for (size_t k = 0; k < dense_terms.size(); ++k) {
        const int32_t t = dense_terms[k];
        const int64_t off = d->term_off[t], n = d->term_off[t + 1] - off;
        ix->term_plane[(size_t)t] = (int32_t)k;
        plane_fill_kernel<<<(unsigned)((n + 255) / 256), 256>>>(ix->post_docs.p + off, ix->post_f8.p + off, n,
                                                                 ix->dense_tf.p + (size_t)k * stride);
      }
      NRT_CUDA_TRY(cudaGetLastError());
      const int64_t n2 = (int64_t)(stride / 4) * (int64_t)dense_terms.size();
      if ((rc = ix->dense_tf2.alloc((size_t)n2))) return rc;
      plane_pack2_kernel<<<(unsigned)((n2 + 255) / 256), 256>>>(ix->dense_tf.p, n2, ix->dense_tf2.p);
      NRT_CUDA_TRY(cudaGetLastError());
    }
  }
  // skip data: posting offsets at every stream-kernel granule boundary for the lists long enough to make the
  // per-batch lower_bound searches expensive (>= 4096 postings); a batch then copies the row instead of searching
  ix->term_gran.assign((size_t)d->n_terms, -1);
  {
    const int32_t n_gran = std::max<int32_t>(1, (int32_t)(((int64_t)d->n_docs + v2::kGran - 1) / v2::kGran));
    std::vector<int64_t> row_off; std::vector<int32_t> row_n;
    const size_t max_rows = (size_t)((2ll << 30) / ((int64_t)(n_gran + 1) * 4));
    for (int32_t t = 0; t < d->n_terms && row_off.size() < max_rows; ++t) {
      const int64_t n = d->term_off[t + 1] - d->term_off[t];
      if (n >= 4096) { ix->term_gran[(size_t)t] = (int32_t)row_off.size(); row_off.push_back(d->term_off[t]); row_n.push_back((int32_t)n); }
    }
    ix->gran_n = n_gran;
    if (!row_off.empty()) {
      DevBuf<int64_t> d_ro; DevBuf<int32_t> d_rn;
      if ((rc = d_ro.upload(row_off.data(), row_off.size()))) return rc;
      if ((rc = d_rn.upload(row_n.data(), row_n.size()))) return rc;
      const int64_t total = (int64_t)row_off.size() * (n_gran + 1);
      if ((rc = ix->gran_tab.alloc((size_t)total))) return rc;
      v2::GranTabLaunch G;
      G.post_docs = ix->post_docs.p; G.row_off = d_ro.p; G.row_n = d_rn.p; G.n_rows = (int32_t)row_off.size();
      G.n_gran = n_gran; G.n_docs = d->n_docs; G.tab = ix->gran_tab.p;
      v2::gran_table_kernel<<<(unsigned)((total + 255) / 256), 256>>>(G);
      NRT_CUDA_TRY(cudaGetLastError());
      NRT_CUDA_TRY(cudaDeviceSynchronize());
    }
  }
  // norms + BM25 caches
  {
    std::vector<const uint8_t*> ptrs((size_t)d->n_fields, nullptr);
    std::vector<float> caches((size_t)d->n_fields * 256);
    std::vector<uint8_t> min_norm((size_t)d->n_fields, 1);
    ix->field_has_norms.resize(d->n_fields);
    for (int f = 0; f < d->n_fields; ++f) {
      ix->norms.emplace_back(new DevBuf<uint8_t>);
      const uint8_t* h = d->norms ? d->norms[f] : nullptr;
      ix->field_has_norms[f] = h != nullptr;
      if (h) {
        if ((rc = ix->norms.back()->upload(h, (size_t)d->n_docs))) return rc;
        ptrs[f] = ix->norms.back()->p;
        int mn = 256;
        for (int32_t i = 0; i < d->n_docs; ++i) if (h[i] != 0 && h[i] < mn) mn = h[i];
        min_norm[f] = (uint8_t)(mn == 256 ? 0 : mn);
      }
      float k1 = d->field_k1 ? d->field_k1[f] : 1.2f, b = d->field_b ? d->field_b[f] : 0.75f;
      ix->field_k1.push_back(k1); ix->field_b.push_back(b);
      int64_t dc = d->field_doc_count[f];
      float avgdl = dc > 0 ? (float)((double)d->field_sum_ttf[f] / (double)dc) : 1.0f;
      bm25_cache(k1, b, avgdl, &caches[(size_t)f * 256]);
    }
    if ((rc = ix->norms_ptrs.upload(ptrs.data(), ptrs.size()))) return rc;
    if ((rc = ix->caches.upload(caches.data(), caches.size()))) return rc;
    if ((rc = ix->field_min_norm.upload(min_norm.data(), min_norm.size()))) return rc;
  }
  // numeric doc-value columns (int32 when the value range allows: 4 B/doc gathers)
  {
    std::vector<const int64_t*> p64((size_t)d->n_columns, nullptr);
    std::vector<const int32_t*> p32((size_t)d->n_columns, nullptr);
    std::vector<const uint8_t*> ph((size_t)d->n_columns, nullptr);
    std::vector<const int64_t*> pmo((size_t)d->n_columns, nullptr), pmv((size_t)d->n_columns, nullptr);
    ix->col_multi.assign((size_t)d->n_columns, 0);
    for (int c = 0; c < d->n_columns; ++c) {
      ix->col64.emplace_back(new DevBuf<int64_t>);
      ix->col32.emplace_back(new DevBuf<int32_t>);
      ix->col_has.emplace_back(new DevBuf<uint8_t>);
      ix->colmv_off.emplace_back(new DevBuf<int64_t>);
      const int64_t* h = d->columns[c];
      if (!h) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: NULL column");
      const int64_t* mo = d->column_offsets ? d->column_offsets[c] : nullptr;
      if (mo) {   // SORTED_NUMERIC: CSR of values per doc
        if (mo[0] != 0) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: column_offsets[c][0] must be 0");
        for (int32_t i = 0; i < d->n_docs; ++i) {
          if (mo[i + 1] < mo[i]) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: column_offsets must be non-decreasing");
          for (int64_t p = mo[i] + 1; p < mo[i + 1]; ++p)
            if (h[p] < h[p - 1]) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: the values of a doc must be ascending (SortedNumericDocValues)");
        }
        ix->col_multi[(size_t)c] = 1;
        if ((rc = ix->colmv_off.back()->upload(mo, (size_t)d->n_docs + 1))) return rc;
        if ((rc = ix->col64.back()->upload(h, (size_t)std::max<int64_t>(mo[d->n_docs], 1)))) return rc;
        pmo[c] = ix->colmv_off.back()->p; pmv[c] = ix->col64.back()->p;
        continue;
      }
      bool fits = true;
      for (int32_t i = 0; i < d->n_docs; ++i) if (h[i] < INT32_MIN || h[i] > INT32_MAX) { fits = false; break; }
      if (fits) {
        std::vector<int32_t> tmp((size_t)d->n_docs);
        for (int32_t i = 0; i < d->n_docs; ++i) tmp[i] = (int32_t)h[i];
        if ((rc = ix->col32.back()->upload(tmp.data(), tmp.size()))) return rc;
        p32[c] = ix->col32.back()->p;
      } else {
        if ((rc = ix->col64.back()->upload(h, (size_t)d->n_docs))) return rc;
        p64[c] = ix->col64.back()->p;
      }
      const uint8_t* hh = d->column_has ? d->column_has[c] : nullptr;
      if (hh) { if ((rc = ix->col_has.back()->upload(hh, (size_t)d->n_docs))) return rc; ph[c] = ix->col_has.back()->p; }
    }
    // sort codes of every column (TopFieldCollector path): one device sort per column at build time
    if (d->n_columns > 0 && d->n_docs > 0) {
      DevBuf<uint64_t> keys; DevBuf<int32_t> idx, rank;
      if ((rc = keys.alloc((size_t)d->n_docs)) || (rc = idx.alloc((size_t)d->n_docs)) || (rc = rank.alloc((size_t)d->n_docs))) return rc;
      for (int c = 0; c < d->n_columns; ++c) {
        ix->col_code.emplace_back(new DevBuf<uint32_t>);
        ix->col_distinct.emplace_back(new DevBuf<uint64_t>);
        if (ix->col_multi[(size_t)c]) { ix->col_n_distinct.push_back(0); continue; }   // no sort / terms on a multi-valued column
        if ((rc = ix->col_code.back()->alloc((size_t)d->n_docs))) return rc;
        if ((rc = ix->col_distinct.back()->alloc((size_t)d->n_docs))) return rc;
        NRT_CUDA_TRY(cudaMemset(ix->col_code.back()->p, 0, ix->col_code.back()->bytes()));
        int32_t nd = 0;
        if ((rc = sort_codes_build(p64[c], p32[c], ph[c], d->n_docs, ix->col_code.back()->p, keys.p, idx.p, rank.p,
                                   ix->col_distinct.back()->p, &nd))) return rc;
        ix->col_n_distinct.push_back(nd);
      }
    }
    if ((rc = ix->col64_ptrs.upload(p64.data(), p64.size()))) return rc;
    if ((rc = ix->col32_ptrs.upload(p32.data(), p32.size()))) return rc;
    if ((rc = ix->col_has_ptrs.upload(ph.data(), ph.size()))) return rc;
    if ((rc = ix->colmv_off_ptrs.upload(pmo.data(), pmo.size()))) return rc;
    if ((rc = ix->colmv_val_ptrs.upload(pmv.data(), pmv.size()))) return rc;
  }
  // index-time impacts (list-wide score bounds for MAXSCORE)
  if ((rc = compute_term_max_x(ix.get()))) return rc;
  if (d->live_docs && (rc = upload_live_docs(ix.get(), d->live_docs))) return rc;
  // vectors
  if (d->vec_dims > 0 && d->vec_count > 0) {
    if (!d->vectors) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: NULL vectors");
    if (d->vec_element_type != NRTGPU_VEC_FLOAT32 && d->vec_element_type != NRTGPU_VEC_INT8) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: bad vec_element_type");
    if (d->vec_dims > 4096) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_build: vector dims > 4096 (VectorFieldDef.java:96)");
    ix->vec_dims = d->vec_dims; ix->vec_sim = d->vec_similarity; ix->vec_count = d->vec_count;
    ix->vec_is_byte = d->vec_element_type == NRTGPU_VEC_INT8;
    if (ix->vec_is_byte) {   // bytes are held as exact floats (every int8 is exact in fp32 and in bf16)
      const int8_t* src = reinterpret_cast<const int8_t*>(d->vectors);
      std::vector<float> tmp((size_t)d->vec_count * d->vec_dims);
      for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = (float)src[i];
      if ((rc = ix->vectors.upload(tmp.data(), tmp.size()))) return rc;
    } else if ((rc = ix->vectors.upload(d->vectors, (size_t)d->vec_count * d->vec_dims))) return rc;
    if (d->vec_docs) { if ((rc = ix->vec_docs.upload(d->vec_docs, (size_t)d->vec_count))) return rc; }
    if ((rc = ix->vec_norm2.alloc((size_t)d->vec_count))) return rc;
    if ((rc = knn_prepare_norms(ix->vectors.p, ix->vec_count, ix->vec_dims, ix->vec_norm2.p))) return rc;
    {
      DevBuf<unsigned int> d_mx;
      if ((rc = d_mx.alloc(1))) return rc;
      NRT_CUDA_TRY(cudaMemset(d_mx.p, 0, sizeof(unsigned int)));
      knn_max_norm2_kernel<<<256, 256>>>(ix->vec_norm2.p, ix->vec_count, d_mx.p);
      NRT_CUDA_TRY(cudaGetLastError());
      float mx = 0.0f;
      NRT_CUDA_TRY(cudaMemcpy(&mx, d_mx.p, sizeof(float), cudaMemcpyDeviceToHost));
      ix->vec_dmax = std::sqrt(mx) * 1.0001f;
    }
    if (d->vec_dims % 8 == 0) {   // TMA needs 16-byte row pitch
      if ((rc = ix->vec_bf16.alloc((size_t)d->vec_count * d->vec_dims))) return rc;
      tc::f32_to_bf16_kernel<<<1024, 256>>>(ix->vectors.p, ix->vec_bf16.p, (size_t)d->vec_count * d->vec_dims);
      NRT_CUDA_TRY(cudaGetLastError());
      if ((rc = tc::make_tensor_map_bf16(&ix->vec_tmap, ix->vec_bf16.p, (uint64_t)d->vec_count, (uint64_t)d->vec_dims, tc::BN))) return rc;
      if ((rc = tc::make_tensor_map_bf16(&ix->vec_tmap128, ix->vec_bf16.p, (uint64_t)d->vec_count, (uint64_t)d->vec_dims, tc::BN3))) return rc;
      if ((rc = ix->vec_ab.alloc((size_t)d->vec_count))) return rc;
      knn_ab_kernel<<<(d->vec_count + 255) / 256, 256>>>(ix->vec_norm2.p, d->vec_count, d->vec_similarity, ix->vec_ab.p);
      NRT_CUDA_TRY(cudaGetLastError());
      ix->vec_tc = true;
    }
  }
  ix->device_bytes = (int64_t)(ix->gran_tab.bytes() + ix->dense_tf.bytes() + ix->dense_tf2.bytes() + ix->post_docs.bytes() + ix->post_f8.bytes() + ix->exc_pos.bytes() + ix->exc_freq.bytes() +
                               ix->caches.bytes() + ix->live_bits.bytes() + ix->vectors.bytes() + ix->vec_norm2.bytes() +
                               ix->vec_docs.bytes() + ix->vec_bf16.bytes());
  for (auto& b : ix->norms) ix->device_bytes += (int64_t)b->bytes();
  for (auto& b : ix->col64) ix->device_bytes += (int64_t)b->bytes();
  for (auto& b : ix->col32) ix->device_bytes += (int64_t)b->bytes();
  for (auto& b : ix->col_has) ix->device_bytes += (int64_t)b->bytes();
  for (auto& b : ix->colmv_off) ix->device_bytes += (int64_t)b->bytes();
  for (auto& b : ix->col_code) ix->device_bytes += (int64_t)b->bytes();
  for (auto& b : ix->col_distinct) ix->device_bytes += (int64_t)b->bytes();
  NRT_CUDA_TRY(cudaDeviceSynchronize());
  *out = ix.release();
  return NRTGPU_OK;
}

int nrtgpu_index_close(nrtgpu_index* ix) {
  if (!ix) return NRTGPU_OK;
  cudaSetDevice(ix->ctx->device);
  delete ix;
  return NRTGPU_OK;
}

int64_t nrtgpu_index_device_bytes(const nrtgpu_index* ix) { return ix ? ix->device_bytes : 0; }

int nrtgpu_index_set_live_docs(nrtgpu_index* ix, const uint8_t* live_docs) {
  if (!ix) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_set_live_docs: NULL index");
  NRT_CUDA_TRY(cudaSetDevice(ix->ctx->device));
  NRT_CUDA_TRY(cudaDeviceSynchronize());   // searches in flight on this image finish against the old bitmap
  return upload_live_docs(ix, live_docs);
}

int nrtgpu_index_update_stats(nrtgpu_index* ix, const int64_t* term_df, const int64_t* field_doc_count, const int64_t* field_sum_ttf) {
  if (!ix || !field_doc_count || !field_sum_ttf) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_index_update_stats: NULL argument");
  NRT_CUDA_TRY(cudaSetDevice(ix->ctx->device));
  NRT_CUDA_TRY(cudaDeviceSynchronize());
  if (term_df) ix->term_df.assign(term_df, term_df + ix->n_terms);
  ix->field_doc_count.assign(field_doc_count, field_doc_count + ix->n_fields);
  ix->field_sum_ttf.assign(field_sum_ttf, field_sum_ttf + ix->n_fields);
  std::vector<float> caches((size_t)ix->n_fields * 256);
  for (int f = 0; f < ix->n_fields; ++f) {
    const int64_t dc = ix->field_doc_count[(size_t)f];
    const float avgdl = dc > 0 ? (float)((double)ix->field_sum_ttf[(size_t)f] / (double)dc) : 1.0f;
    bm25_cache(ix->field_k1[(size_t)f], ix->field_b[(size_t)f], avgdl, &caches[(size_t)f * 256]);
  }
  int rc;
  if ((rc = ix->caches.upload(caches.data(), caches.size()))) return rc;
  return compute_term_max_x(ix);   // the impacts are functions of the length cache
}

// ---- batch compilation: flat BooleanQuery -> DevQuery/DevClause, driver selection, work list ----
// compile + upload a batch into `b` (buffers are reused when large enough); asynchronous on `st`
static int batch_build(nrtgpu_batch* b, nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                       const nrtgpu_query* queries, int32_t nq, int32_t top_k, int32_t total_hits_threshold,
                       int32_t flags, cudaStream_t st, const nrtgpu_sort* sort = nullptr, const nrtgpu_aggregation* aggs = nullptr,
                       int32_t n_aggs = 0) {
  if (!ix || !queries || (n_clauses > 0 && !clauses)) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_batch_prepare: NULL argument");
  b->aggs.clear();
  if (n_aggs > 0) {
    if (!aggs || n_aggs > kMaxAggs) NRT_FAIL(NRTGPU_ERR_INVALID, "at most 8 aggregations per search");
    if (ix && ix->ctx->engine_stream) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "aggregations need the probe engine");
    for (int i = 0; i < n_aggs; ++i) {
      const nrtgpu_aggregation& a = aggs[i];
      if (a.kind < NRTGPU_AGG_TERMS || a.kind > NRTGPU_AGG_SUM) NRT_FAIL(NRTGPU_ERR_INVALID, "bad aggregation kind");
      if (!ix || a.column < 0 || a.column >= ix->n_columns) NRT_FAIL(NRTGPU_ERR_INVALID, "aggregation column out of range");
      if (a.value_type < 0 || a.value_type > 2) NRT_FAIL(NRTGPU_ERR_INVALID, "bad aggregation value_type");
      if (ix->col_multi[(size_t)a.column]) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "aggregation on a multi-valued column");
      if (a.kind == NRTGPU_AGG_TERMS) {
        if (a.size <= 0 || a.size > kAggChunk) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "terms aggregation: size must be in [1, 2048]");
        const int64_t cells = (int64_t)nq * ix->col_n_distinct[(size_t)a.column];
        if (cells * 4 > (2ll << 30)) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "terms aggregation: batch x distinct values exceeds the 2 GB count table");
      }
      b->aggs.push_back(a);
    }
    total_hits_threshold = INT32_MAX;   // RelevanceCollector.java:55-62: additional collectors force exact collection
  }
  const bool sorted = sort && sort->kind != NRTGPU_SORT_RELEVANCE;
  b->sort_kind = sorted ? sort->kind : 0; b->sort_column = sorted ? sort->column : 0; b->sort_reverse = sorted ? (sort->reverse != 0) : 0;
  b->sort_missing_value = sorted ? sort->missing_value : 0;
  if (sorted) {
    if (sort->kind != NRTGPU_SORT_COLUMN && sort->kind != NRTGPU_SORT_DOCID) NRT_FAIL(NRTGPU_ERR_INVALID, "bad sort kind");
    if (sort->kind == NRTGPU_SORT_COLUMN && (sort->column < 0 || sort->column >= ix->n_columns))
      NRT_FAIL(NRTGPU_ERR_INVALID, "sort column out of range (field does not support sorting: no doc values)");
    if (sort->kind == NRTGPU_SORT_COLUMN && ix->col_multi[(size_t)sort->column]) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "sort on a multi-valued column");
    if (ix->ctx->engine_stream) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "sorted search needs the probe engine");
    total_hits_threshold = INT32_MAX;   // every match is visited: exact totalHits
  }
  if (nq <= 0) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_batch_prepare: nq must be > 0");
  // LazyQueueTopScoreDocCollectorManager.java:93-96: numHits must be > 0
  if (top_k <= 0) NRT_FAIL(NRTGPU_ERR_INVALID, "numHits must be > 0; please use TotalHitCountCollectorManager if you just need the total hit count");
  if (top_k > kMaxTopK) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "nrtgpu_batch_prepare: top_k > 1024 is not on the GPU path");
  if (total_hits_threshold < 0) NRT_FAIL(NRTGPU_ERR_INVALID, "totalHitsThreshold must be >= 0");
  NRT_CUDA_TRY(cudaSetDevice(ix->ctx->device));
  b->ix = ix; b->nq = nq; b->top_k = top_k;
  b->alg_postings = 0; b->ran = false; b->runs_recorded = 0;
  // LazyQueueTopScoreDocCollectorManager.java:102: totalHitsThreshold = max(totalHitsThreshold, numHits);
  // Integer.MAX_VALUE <=> ScoreMode.COMPLETE (LazyQueueTopScoreDocCollector.java:68-70): exact counts, no list skipping
  b->threshold = (total_hits_threshold == INT32_MAX || (flags & NRTGPU_FLAG_NO_PRUNING))
                     ? (int64_t)INT32_MAX : (int64_t)std::max(total_hits_threshold, top_k);
  b->exhaustive = b->threshold == (int64_t)INT32_MAX;
  // slice size depends on the kernel: decided after the clauses are known (wide queries / large top_k -> bool_window_kernel)
  int64_t slice_docs = (int64_t)kSliceWindows * kWindowDocs;
  std::vector<DevClause> dc;
  std::vector<DevQuery> dq((size_t)nq);
  dc.reserve((size_t)n_clauses);
  int max_terms = 0;
  for (int qi = 0; qi < nq; ++qi) {
    const nrtgpu_query& q = queries[qi];
    if (q.clause_begin < 0 || q.clause_end < q.clause_begin || q.clause_end > n_clauses)
      NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_batch_prepare: clause range out of bounds");
    if (q.min_should_match < 0) NRT_FAIL(NRTGPU_ERR_INVALID, "minimumNumberShouldMatch must be >= 0");
    int ncl = q.clause_end - q.clause_begin;
    if (ncl > kMaxClauses) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "nrtgpu_batch_prepare: more than 16 clauses in one BooleanQuery");
    DevQuery& o = dq[(size_t)qi];
    std::memset(&o, 0, sizeof(o));
    o.clause_begin = (int32_t)dc.size(); o.n_clauses = ncl; o.msm = q.min_should_match;
    int n_term = 0, n_req = 0, n_should = 0, n_req_term = 0, n_req_nonterm = 0, n_should_nonterm = 0;
    int best_req_slot = -1; int32_t best_req_n = INT32_MAX;
    uint32_t should_term_mask = 0;
    int field0 = -2;   // -2: no term yet, -1: mixed
    for (int ci = q.clause_begin; ci < q.clause_end; ++ci) {
      const nrtgpu_clause& c = clauses[ci];
      if (c.occur < NRTGPU_SHOULD || c.occur > NRTGPU_MUST_NOT) NRT_FAIL(NRTGPU_ERR_INVALID, "bad occur");
      // QueryNodeMapper.java:127-133: the reference rejects boost < 0 with exactly this message and treats the proto default 0 as
      // "no boost"; the adaptor folds that rule before it fills the clause, so 0 here is a weight of 0, not an error
      if (c.boost < 0.0f) NRT_FAIL(NRTGPU_ERR_INVALID, "Boost must be a positive number");
      DevClause x; std::memset(&x, 0, sizeof(x));
      x.occur = c.occur; x.kind = c.kind; x.slot = -1; x.plane = -1; x.gran_row = -1; x.lo = c.lo; x.hi = c.hi;
      x.scoring = (c.occur == NRTGPU_MUST || c.occur == NRTGPU_SHOULD) ? 1 : 0;
      bool required = (c.occur == NRTGPU_MUST || c.occur == NRTGPU_FILTER);
      if (c.kind == NRTGPU_TERM) {
        if (c.id < 0 || c.id >= ix->n_terms) NRT_FAIL(NRTGPU_ERR_INVALID, "term id out of range");
        if (n_term >= kMaxTermSlots) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "nrtgpu_batch_prepare: more than 8 term clauses in one BooleanQuery");
        int f = ix->term_field[c.id];
        x.post_base = ix->term_off[c.id];
        x.n_post = (int32_t)(ix->term_off[c.id + 1] - ix->term_off[c.id]);
        x.slot = n_term; x.field = f; x.plane = ix->term_plane[c.id]; x.gran_row = ix->term_gran[c.id];
        int64_t df = ix->term_df[c.id];
        // BM25Scorer: weight = boost * idf
        x.weight = c.boost * bm25_idf(df > 0 ? df : 1, ix->field_doc_count[f]);
        {
          volatile float t1 = 1.0f + ix->term_max_x[c.id];
          volatile float t2 = x.weight / t1;
          x.ub = x.weight - t2;
        }
        if (required) { o.req_term_mask |= 1u << n_term; ++n_req_term; if (x.n_post < best_req_n) { best_req_n = x.n_post; best_req_slot = n_term; } }
        if (c.occur == NRTGPU_MUST_NOT) o.not_term_mask |= 1u << n_term;
        if (c.occur == NRTGPU_SHOULD) should_term_mask |= 1u << n_term;
        if (c.occur == NRTGPU_MUST) o.must_term_mask |= 1u << n_term;
        if (x.scoring) field0 = (field0 == -2 || field0 == f) ? f : -1;
        if (x.scoring) b->alg_postings += x.n_post; else b->alg_postings += x.n_post;
        ++n_term;
      } else if (c.kind == NRTGPU_RANGE_I64) {
        if (c.id < 0 || c.id >= ix->n_columns) NRT_FAIL(NRTGPU_ERR_INVALID, "column id out of range");
        x.col = c.id; x.weight = c.boost;  // constant-score query: score = boost
        o.has_nonterm = 1;
        if (x.scoring) o.nonterm_scoring = 1;
        if (required) ++n_req_nonterm;
        if (c.occur == NRTGPU_SHOULD) ++n_should_nonterm;
      } else if (c.kind == NRTGPU_MATCH_ALL) {
        x.weight = c.boost;
        o.has_nonterm = 1;
        if (x.scoring) o.nonterm_scoring = 1;
        if (required) ++n_req_nonterm;
        if (c.occur == NRTGPU_SHOULD) ++n_should_nonterm;
      } else NRT_FAIL(NRTGPU_ERR_INVALID, "bad clause kind");
      if (required) ++n_req;
      if (c.occur == NRTGPU_SHOULD) ++n_should;
      dc.push_back(x);
    }
    o.n_term = n_term; o.n_req = n_req;
    o.should_term_mask = should_term_mask;
    o.single_field = field0 == -2 ? 0 : field0;
    o.need_should = q.min_should_match > 0 ? q.min_should_match : (n_req == 0 ? 1 : 0);
    max_terms = std::max(max_terms, n_term);
    if (q.min_should_match > n_should || (n_req == 0 && n_should == 0)) o.empty = 1;
    // driver selection
    if (n_req_term > 0) o.driver_mask = 1u << best_req_slot;           // rarest required posting list leads
    else if (n_req_nonterm > 0 || n_should_nonterm > 0) o.dense_driver = 1;  // no posting list can lead
    else o.driver_mask = should_term_mask;                             // pure disjunction: every SHOULD list drives
    uint32_t all_terms = n_term >= 32 ? 0xffffffffu : ((1u << n_term) - 1u);
    o.has_non_driver = (!o.dense_driver && (all_terms & ~o.driver_mask)) ? 1 : 0;
    if (q.has_after && sorted) o.has_after = 1;   // after_key is patched on the device (sort_after_kernel)
    else if (q.has_after) {
      o.has_after = 1;
      int64_t local = (int64_t)q.after_doc - ix->doc_base;
      uint32_t ord = float_to_ordered(q.after_score);
      if (local < 0) o.after_key = ((uint64_t)ord + 1ull) << 32;             // every doc here follows afterDoc
      else if (local >= ix->n_docs) o.after_key = make_key(q.after_score, INT32_MAX);  // every doc here precedes it
      else o.after_key = make_key(q.after_score, (int32_t)local);
    }
  }
  b->wide_slots = max_terms > 4 || top_k > v2::kMaxTopKStream;
  b->use_probe = !b->wide_slots && !ix->ctx->engine_stream;
  if (n_aggs > 0 && b->wide_slots) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "aggregations: more than 4 term clauses or top_k > 512 is not on the GPU path");
  if (sorted && b->wide_slots) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "sorted search: more than 4 term clauses or top_k > 512 is not on the GPU path");
  if (!b->wide_slots) {
    // slices of equal size, a multiple of the 1024-doc granule, at most 512K docs: a 1.25M-doc shard is 3 x 417K, not 2.38 -> 3 x 512K
    // (probe kernel: slices of up to slice_gran granules -- NRTGPU_SLICE_GRAN, default and maximum 512: the MAXSCORE roles of an
    //  item are fixed when it starts, so larger slices prune with staler thresholds -- measured slower -- and smaller ones pay
    //  the per-item set-up more often)
    const int64_t max_slice_docs = (!ix->ctx->engine_stream) ? (int64_t)ix->ctx->slice_gran * v2::kGran : (int64_t)v2::kSliceDocs;
    const int64_t n_sl = std::max<int64_t>(1, ((int64_t)ix->n_docs + max_slice_docs - 1) / max_slice_docs);
    slice_docs = (((int64_t)ix->n_docs + n_sl - 1) / n_sl + v2::kGran - 1) / v2::kGran * v2::kGran;
    if (slice_docs < v2::kGran) slice_docs = v2::kGran;
  }
  b->slice_docs = (int32_t)slice_docs;
  b->n_slices = (int32_t)std::max<int64_t>(1, ((int64_t)ix->n_docs + slice_docs - 1) / slice_docs);
  // work list, slice-major so that concurrently resident CTAs share postings of the same doc range in L2;
  // inside a slice, longer queries first
  std::vector<int32_t> order;
  std::vector<int64_t> cost((size_t)nq, 0);
  for (int qi = 0; qi < nq; ++qi) {
    if (dq[qi].empty) continue;
    order.push_back(qi);
    for (int c = 0; c < dq[qi].n_clauses; ++c) cost[qi] += dc[(size_t)dq[qi].clause_begin + c].n_post;
    if (dq[qi].dense_driver) cost[qi] += ix->n_docs;
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return cost[a] > cost[c]; });
  if (b->use_probe && !ix->ctx->order_by_cost) {
    // probe kernel: inside a slice, queries that share their densest tf plane are adjacent in the queue, so the plane's
    // bytes of the slice are gathered by CTAs that run together and stay in L2 between them (longer queries first inside
    // a cluster, clusters of the densest -- most shared -- planes first)
    std::vector<int64_t> ckey((size_t)nq, INT64_MAX);
    for (int qi : order) {
      int64_t best_n = -1;
      for (int c = 0; c < dq[qi].n_clauses; ++c) {
        const DevClause& x = dc[(size_t)dq[qi].clause_begin + c];
        if (x.kind == NRTGPU_TERM && x.plane >= 0 && x.n_post > best_n) { best_n = x.n_post; ckey[(size_t)qi] = ((int64_t)(INT32_MAX - x.n_post) << 32) | (uint32_t)x.plane; }
      }
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return ckey[(size_t)a] < ckey[(size_t)c]; });
  }
  // pure disjunctions of scoring term clauses over one text field (no deletes) run in their own kernel instantiation
  // (tf-pattern bound, deferred scoring, MAXSCORE): their work items come first
  auto is_simple = [&](int qi) {
    const DevQuery& o = dq[(size_t)qi];
    return !sorted && n_aggs == 0 && o.single_field >= 0 && !o.has_nonterm && !o.nonterm_scoring && (b->use_probe || ix->live_bits.p == nullptr) && o.n_req == 0 &&
           o.not_term_mask == 0 && o.msm <= 1 && !o.dense_driver;
  };
  // 0: probe kernel, simple; 1: probe kernel, generic (a posting list leads); 2: window/stream kernel (no list can lead)
  auto engine_class = [&](int qi) {
    if (is_simple(qi)) return 0;
    if (!b->use_probe) return 2;
    return 1;   // the probe kernel also sweeps whole doc ranges when no posting list can lead (match-all / range-led queries)
  };
  std::vector<int32_t>& wq = b->h_wq; std::vector<int32_t>& ws = b->h_ws;
  wq.clear(); ws.clear();
  // warm-up items (large shards): a query first sweeps the leading kWarmGran granules of slice 0 as a work item of its
  // own, ahead of everything else, so that its other work items start with a threshold and a hit count (MAXSCORE can
  // prune from the first slice on). Window kernel: TOP_SCORES mode, queries with a dense list; probe kernel: both score
  // modes, every query whose lists are expected to hold 2 * top_k hits in those granules.
  const bool warm_ok = !b->wide_slots && (b->use_probe || b->threshold < (int64_t)INT32_MAX) &&
                       (int64_t)ix->n_docs >= (int64_t)ix->ctx->warm_min_docs;
  std::vector<uint8_t> has_warm((size_t)nq, 0);
  // probe kernel, pure disjunctions on a shard without deletes: the longest list is a lower bound of the matching docs
  // (pruning may start as soon as that exceeds totalHitsThreshold), and the warm-up SWEEPS the first 32K postings of the
  // highest-bound (rarest) list over the whole shard instead of the first 32K docs (flags 4; NRTGPU_WARM=docs: old style)
  b->h_known.assign((size_t)nq, 0ull);
  const bool limits_may_apply = true;   // (terminateAfter counts collected hits only: the kernel never uses known_hits for it)
  (void)limits_may_apply;
  if (b->use_probe && !ix->live_bits.p && !sorted && n_aggs == 0)
    for (int qi : order) {
      if (!is_simple(qi)) continue;
      int64_t mx = 0;
      for (int c = 0; c < dq[qi].n_clauses; ++c) mx = std::max<int64_t>(mx, dc[(size_t)dq[qi].clause_begin + c].n_post);
      b->h_known[(size_t)qi] = (unsigned long long)mx;
    }
  if (warm_ok) {
    for (int qi : order) {
      if (!is_simple(qi)) continue;
      // (not with searchAfter: a doc whose LOWER-BOUND key passes the after filter may in truth lie on an earlier page, and
      //  would be counted towards the k keys that justify the threshold)
      if (b->use_probe && ix->ctx->warm_sweep && !dq[qi].has_after) {
        int best = -1; float best_ub = -1.0f;
        for (int c = 0; c < dq[qi].n_clauses; ++c) {
          const DevClause& x = dc[(size_t)dq[qi].clause_begin + c];
          if (x.kind == NRTGPU_TERM && x.ub > best_ub) { best_ub = x.ub; best = c; }
        }
        if (best >= 0 && dc[(size_t)dq[qi].clause_begin + best].n_post >= 2 * (int64_t)top_k) {
          wq.push_back(qi); ws.push_back(0 | (dc[(size_t)dq[qi].clause_begin + best].slot << 16) | (4 << 24));
          continue;   // (has_warm stays 0: the query's slice-0 items cover the whole slice)
        }
      }
      if (b->use_probe) {
        if (cost[qi] * (int64_t)(v2::kWarmGran * v2::kGran) >= 2ll * top_k * (int64_t)ix->n_docs) has_warm[(size_t)qi] = 1;
      } else {
        for (int c = 0; c < dq[qi].n_clauses; ++c) {
          const DevClause& x = dc[(size_t)dq[qi].clause_begin + c];
          if (x.kind == NRTGPU_TERM && (int64_t)x.n_post * 64 >= (int64_t)ix->n_docs) has_warm[(size_t)qi] = 1;
        }
      }
      if (has_warm[(size_t)qi]) { wq.push_back(qi); ws.push_back(0 | (1 << 24)); }
    }
  }
  // heavy (query, slice) pairs are split into 2..16 parts of equal granule ranges, so that no single work item is a
  // large share of the launch (the longest item bounds the kernel time from below: what limits small shards)
  const int gran_per_slice = (int)(slice_docs / v2::kGran);
  const int n_gran_h = std::max<int>(1, (int)(((int64_t)ix->n_docs + v2::kGran - 1) / v2::kGran));
  std::vector<uint8_t> lparts((size_t)nq, 0);
  int lp_max = 0;
  if (b->use_probe) {
    int64_t total_cost = 0;
    for (int qi : order) total_cost += cost[qi];
    const int64_t per_cta = total_cost / ((int64_t)v3::kCtasA * ix->ctx->sm_count);
    const int64_t item_max_top = std::max<int64_t>(ix->ctx->item_postings, per_cta / ix->ctx->item_share);
    const int64_t item_max_full = std::max<int64_t>(ix->ctx->item_postings, per_cta / ix->ctx->item_share_full);
    for (int qi : order) {
      // pruned sweeps (TOP_SCORES pure disjunctions) skip most of a heavy item's postings; launches that visit every
      // posting are split finer (their longest item is the tail of the launch)
      const int64_t item_max = (is_simple(qi) && b->threshold < (int64_t)INT32_MAX) ? item_max_top : item_max_full;
      const int64_t per_slice = cost[qi] / std::max(1, (int)b->n_slices);
      int lp = 0;
      while (lp < 4 && (per_slice >> lp) > item_max && (gran_per_slice >> (lp + 1)) >= 8) ++lp;
      lparts[(size_t)qi] = (uint8_t)lp;
      lp_max = std::max(lp_max, lp);
    }
  }
  b->parts_max = 1 << lp_max;
  b->n_lists = b->n_slices * b->parts_max + (warm_ok ? 1 : 0);
  auto push_items = [&](int qi, int s) {   // the parts of (query, slice) that hold at least one granule
    const int lp = lparts[(size_t)qi], P = 1 << lp;
    const int g_count = std::min(gran_per_slice, n_gran_h - s * gran_per_slice);
    const int fine = (gran_per_slice + b->parts_max - 1) / b->parts_max, kfine = b->parts_max >> lp;   // as the kernel decodes them
    const bool behind_warm = s == 0 && has_warm[(size_t)qi];
    for (int p = 0; p < P; ++p) {
      int lo = std::min(g_count, p * kfine * fine), hi = (p + 1) * kfine >= b->parts_max ? g_count : std::min(g_count, (p + 1) * kfine * fine);
      if (behind_warm) lo = std::max(lo, std::min(g_count, (int)v2::kWarmGran));
      if (P > 1 && lo >= hi) continue;
      wq.push_back(qi); ws.push_back(s | (p << 16) | (lp << 20) | (behind_warm ? (2 << 24) : 0));
    }
  };
  int32_t class_end[3] = {0, 0, 0};
  for (int cls = 0; cls < 3; ++cls) {
    if (ix->ctx->order_lpt && b->use_probe) {   // longest query first, its slices together (experiment: NRTGPU_ORDER=lpt)
      for (int qi : order) if (engine_class(qi) == cls)
        for (int s = 0; s < b->n_slices; ++s) push_items(qi, s);
    } else {
      for (int s = 0; s < b->n_slices; ++s)
        for (int qi : order) if (engine_class(qi) == cls) push_items(qi, s);
    }
    class_end[cls] = (int32_t)wq.size();
  }
  b->n_work = (int32_t)wq.size();
  b->n_work_simple = class_end[0];
  if (b->use_probe) { b->n_probe_simple = class_end[0]; b->n_probe_generic = class_end[1] - class_end[0]; b->n_stream = class_end[2] - class_end[1]; }
  else { b->n_probe_simple = b->n_probe_generic = 0; b->n_stream = b->wide_slots ? 0 : b->n_work; }
  b->h_dc.swap(dc); b->h_dq.swap(dq);   // kept alive until the next compilation: the uploads below are asynchronous
  int rc;
  if ((rc = b->clauses.upload_async(b->h_dc.data(), b->h_dc.size(), st))) return rc;
  if ((rc = b->queries.upload_async(b->h_dq.data(), b->h_dq.size(), st))) return rc;
  if ((rc = b->work_query.upload_async(wq.data(), wq.size(), st))) return rc;
  if ((rc = b->work_slice.upload_async(ws.data(), ws.size(), st))) return rc;
  if ((rc = b->known_hits.upload_async(b->h_known.data(), b->h_known.size(), st))) return rc;
  if (sorted) {
    bool any_after = false;
    b->h_after_docs.assign((size_t)nq, 0);
    for (int qi = 0; qi < nq; ++qi) if (queries[qi].has_after) { any_after = true; b->h_after_docs[(size_t)qi] = queries[qi].after_doc; }
    if (any_after && sort->kind == NRTGPU_SORT_COLUMN && !sort->after_values) NRT_FAIL(NRTGPU_ERR_INVALID, "sorted searchAfter needs after_values");
    if ((rc = b->sort_missing_code.alloc(1))) return rc;
    if ((rc = b->after_docs.upload_async(b->h_after_docs.data(), (size_t)nq, st))) return rc;
    if (sort->after_values) { if ((rc = b->after_values.upload_async(sort->after_values, (size_t)nq, st))) return rc; }
    else if ((rc = b->after_values.alloc((size_t)nq))) return rc;
    SortAfterLaunch A;
    A.queries = b->queries.p; A.nq = nq; A.after_docs = b->after_docs.p; A.after_values = b->after_values.p;
    A.kind = sort->kind; A.reverse = sort->reverse != 0; A.doc_base = ix->doc_base; A.n_docs = ix->n_docs;
    const bool col = sort->kind == NRTGPU_SORT_COLUMN;
    A.distinct = col ? ix->col_distinct[(size_t)sort->column]->p : nullptr;
    A.n_distinct = col ? ix->col_n_distinct[(size_t)sort->column] : 0;
    A.missing_value = sort->missing_value; A.missing_code = b->sort_missing_code.p;
    sort_after_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(A);
    NRT_CUDA_TRY(cudaGetLastError());
    if ((rc = b->out_sort_values.alloc((size_t)nq * top_k))) return rc;
  }
  if ((rc = b->theta.alloc((size_t)nq))) return rc;
  if ((rc = b->total_hits.alloc((size_t)nq))) return rc;
  if ((rc = b->slice_keys.alloc((size_t)nq * b->n_lists * top_k))) return rc;
  if ((rc = b->slice_cnt.alloc((size_t)nq * b->n_lists))) return rc;
  if ((rc = b->out_docs.alloc((size_t)nq * top_k))) return rc;
  if ((rc = b->out_scores.alloc((size_t)nq * top_k))) return rc;
  if ((rc = b->out_counts.alloc((size_t)nq))) return rc;
  if ((rc = b->pruned.alloc((size_t)nq))) return rc;
  if ((rc = b->terminated.alloc((size_t)nq))) return rc;
  if (!b->wide_slots) {
    b->n_gran = (int32_t)(((int64_t)ix->n_docs + v2::kGran - 1) / v2::kGran);
    if (b->n_gran < 1) b->n_gran = 1;
    if (b->n_probe_simple + b->n_probe_generic > 0) {
      // probe kernel: posting offsets of every (query, term slot) at the slice boundaries only (skip data for the long
      // lists, one lower_bound for the short ones); the granule offsets inside a slice are read from gran_tab by the kernel
      const int64_t total = (int64_t)nq * v3::kT * ((int64_t)b->n_slices * b->parts_max + 2);
      if ((rc = b->sbounds.alloc((size_t)total))) return rc;
      if ((rc = b->work_counter.alloc(2))) return rc;
      v3::SliceBoundsLaunch S;
      S.ix = ix->view(); S.clauses = b->clauses.p; S.queries = b->queries.p; S.nq = nq; S.n_slices = b->n_slices;
      S.slice_gran = b->slice_docs / v2::kGran; S.n_gran = b->n_gran; S.parts_max = b->parts_max; S.sbounds = b->sbounds.p;
      v3::slice_bounds_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(S);
      NRT_CUDA_TRY(cudaGetLastError());
    }
    if (b->n_stream > 0) {
      // the window/stream kernel reads per-granule posting bounds of every (query, term clause) and per-query score tables
      const int64_t total = (int64_t)nq * v2::kT * (b->n_gran + 1);
      if ((rc = b->gbounds.alloc((size_t)total))) return rc;
      v2::BoundsLaunch B;
      B.ix = ix->view(); B.clauses = b->clauses.p; B.queries = b->queries.p; B.nq = nq; B.n_gran = b->n_gran;
      B.gbounds = b->gbounds.p;
      v2::granule_bounds_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(B);
      NRT_CUDA_TRY(cudaGetLastError());
      if ((rc = b->qtables.alloc((size_t)nq * v2::kQTabFloats))) return rc;
      v2::QTabLaunch Q;
      Q.ix = B.ix; Q.clauses = B.clauses; Q.queries = B.queries; Q.field_min_norm = ix->field_min_norm.p; Q.nq = nq;
      Q.qtables = b->qtables.p;
      if (nq > 0) v2::query_tables_kernel<<<(unsigned)nq, 256, 0, st>>>(Q);
      NRT_CUDA_TRY(cudaGetLastError());
    }
  }
  if (!b->ev[0][0]) for (auto& r : b->ev) for (auto& e : r) NRT_CUDA_TRY(cudaEventCreate(&e));
  return NRTGPU_OK;
}

int nrtgpu_batch_prepare(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                         const nrtgpu_query* queries, int32_t nq, int32_t top_k,
                         int32_t total_hits_threshold, int32_t flags, nrtgpu_batch** out) {
  if (!out) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_batch_prepare: NULL argument");
  std::unique_ptr<nrtgpu_batch> b(new nrtgpu_batch);
  int rc = batch_build(b.get(), ix, clauses, n_clauses, queries, nq, top_k, total_hits_threshold, flags, (cudaStream_t)0);
  if (rc) return rc;
  NRT_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)0));
  *out = b.release();
  return NRTGPU_OK;
}

int nrtgpu_batch_run(nrtgpu_batch* b, void* stream_) {
  if (!b) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_batch_run: NULL batch");
  cudaStream_t st = (cudaStream_t)stream_;
  int rc_dbg = 0;
  NRT_CUDA_TRY(cudaSetDevice(b->ix->ctx->device));
  static const bool keep_theta = getenv("NRTGPU_EXPERIMENT_KEEP_THETA") != nullptr;   // timing experiment only: a run starts with the previous run's thresholds
  if (!keep_theta || !b->ran) {
    NRT_CUDA_TRY(cudaMemsetAsync(b->theta.p, 0, b->theta.bytes(), st));
    NRT_CUDA_TRY(cudaMemsetAsync(b->total_hits.p, 0, b->total_hits.bytes(), st));
  }
  NRT_CUDA_TRY(cudaMemsetAsync(b->slice_cnt.p, 0, b->slice_cnt.bytes(), st));
  NRT_CUDA_TRY(cudaMemsetAsync(b->pruned.p, 0, b->pruned.bytes(), st));
  NRT_CUDA_TRY(cudaMemsetAsync(b->terminated.p, 0, b->terminated.bytes(), st));
  if (b->work_counter.p) NRT_CUDA_TRY(cudaMemsetAsync(b->work_counter.p, 0, b->work_counter.bytes(), st));
  const bool debug = b->ix->ctx->debug_modes;
  cudaEvent_t* ev = b->ev[b->runs_recorded % nrtgpu_batch::kEvRing];
  NRT_CUDA_TRY(cudaEventRecord(ev[0], st));
  if (b->n_work > 0) {
    BoolLaunch L;
    L.ix = b->ix->view();
    L.clauses = b->clauses.p; L.queries = b->queries.p;
    L.work_query = b->work_query.p; L.work_slice = b->work_slice.p;
    L.n_work = b->n_work; L.n_slices = b->n_lists; L.top_k = b->top_k;
    L.theta = b->theta.p; L.total_hits = b->total_hits.p;
    L.slice_keys = b->slice_keys.p; L.slice_cnt = b->slice_cnt.p;
    if (!b->wide_slots) {
      const int n_probe = b->n_probe_simple + b->n_probe_generic;
      if (n_probe > 0) {
        v3::ProbeLaunch P;
        P.ix = L.ix; P.clauses = L.clauses; P.queries = L.queries; P.sbounds = b->sbounds.p;
        P.field_min_norm = b->ix->field_min_norm.p; P.stats = nullptr;
        P.known_hits = b->ix->live_bits.p ? nullptr : b->known_hits.p;   // (deletes installed after the batch was prepared: list lengths no longer bound the hits)
#ifdef NRT_PROBE_KNOCK
        { const char* e = getenv("NRTGPU_KNOCK"); P.knock = e ? atoi(e) : 0; }   // profiling builds only (tools/knock.py)
#else
        P.knock = 0;
#endif
        P.n_lists = b->n_lists; P.parts_max = b->parts_max; P.n_slices = b->n_slices; P.top_k = b->top_k; P.slice_docs = b->slice_docs; P.n_gran = b->n_gran;
        P.threshold = b->threshold; P.pruned = b->pruned.p; P.theta = L.theta; P.total_hits = L.total_hits;
        P.slice_keys = L.slice_keys; P.slice_cnt = L.slice_cnt;
        P.deadline_ns = b->limits_active ? b->deadline_ns : 0; P.clock0 = b->clock0.p; P.timed_out = b->timed_out.p;
        P.terminate_after = b->ta_scalar; P.terminated = b->terminated.p;
        P.sort_kind = b->sort_kind; P.sort_reverse = b->sort_reverse;
        P.sort_codes = b->sort_kind == NRTGPU_SORT_COLUMN ? b->ix->col_code[(size_t)b->sort_column]->p : nullptr;
        P.sort_missing_code = b->sort_missing_code.p;
        P.aggs = nullptr;
        if (!b->aggs.empty()) {
          AggLaunch A; std::memset(&A, 0, sizeof(A));
          A.n_aggs = (int32_t)b->aggs.size();
          for (int i = 0; i < A.n_aggs; ++i) {
            const nrtgpu_aggregation& a = b->aggs[(size_t)i];
            A.a[i].kind = a.kind; A.a[i].column = a.column; A.a[i].value_type = a.value_type;
            if (a.kind == NRTGPU_AGG_TERMS) {
              const int32_t nb = b->ix->col_n_distinct[(size_t)a.column];
              A.a[i].n_buckets = nb;
              if ((rc_dbg = b->agg_counts[i].alloc((size_t)b->nq * (size_t)std::max(nb, 1)))) return rc_dbg;
              NRT_CUDA_TRY(cudaMemsetAsync(b->agg_counts[i].p, 0, b->agg_counts[i].bytes(), st));
              A.a[i].counts = b->agg_counts[i].p; A.codes[i] = b->ix->col_code[(size_t)a.column]->p;
            } else {
              if ((rc_dbg = b->agg_dvals[i].alloc((size_t)b->nq))) return rc_dbg;
              // min starts at +inf / max at -inf in ordered-double space, sum at 0.0 (the "unset" values are applied at fetch)
              const int fill = a.kind == NRTGPU_AGG_MIN ? 0xff : 0x00;
              NRT_CUDA_TRY(cudaMemsetAsync(b->agg_dvals[i].p, fill, b->agg_dvals[i].bytes(), st));
              A.a[i].dvals = b->agg_dvals[i].p;
            }
          }
          if ((rc_dbg = b->agg_launch.upload_async(&A, 1, st))) return rc_dbg;
          NRT_CUDA_TRY(cudaStreamSynchronize(st));   // A is a stack object
          P.aggs = b->agg_launch.p;
        }
        if (P.deadline_ns) {
          NRT_CUDA_TRY(cudaMemsetAsync(b->clock0.p, 0, sizeof(unsigned long long), st));
          NRT_CUDA_TRY(cudaMemsetAsync(b->timed_out.p, 0, b->timed_out.bytes(), st));
        }
        if (debug) {
          if (!b->probe_stats.p && (rc_dbg = b->probe_stats.alloc(32))) return rc_dbg;
          NRT_CUDA_TRY(cudaMemsetAsync(b->probe_stats.p, 0, 32 * sizeof(unsigned long long), st));
        }
        // configuration A (3 CTAs / SM) for the pruned sweeps of TOP_SCORES, B (4 CTAs / SM) where every posting is visited
        const bool cfg_b_simple = ix_ctx_probe_cfg(b->ix->ctx, P.threshold >= (int64_t)INT32_MAX);
        const bool cfg_b_generic = ix_ctx_probe_cfg(b->ix->ctx, true);
        auto launch = [&](auto simple_tag, bool cfg_b, int n_items) {
          constexpr bool S = decltype(simple_tag)::value;
          if (cfg_b) {
            const int grid = std::min(v3::kCtasB * b->ix->ctx->sm_count, n_items);
            if (debug) v3::posting_probe_kernel<S, true, v3::kCtasB, v3::kStageB><<<grid, v3::kThreads, sizeof(v3::ProbeSmemT<v3::kStageB>), st>>>(P);
            else v3::posting_probe_kernel<S, false, v3::kCtasB, v3::kStageB><<<grid, v3::kThreads, sizeof(v3::ProbeSmemT<v3::kStageB>), st>>>(P);
          } else {
            const int grid = std::min(v3::kCtasA * b->ix->ctx->sm_count, n_items);
            if (debug) v3::posting_probe_kernel<S, true, v3::kCtasA, v3::kStageA><<<grid, v3::kThreads, sizeof(v3::ProbeSmemT<v3::kStageA>), st>>>(P);
            else v3::posting_probe_kernel<S, false, v3::kCtasA, v3::kStageA><<<grid, v3::kThreads, sizeof(v3::ProbeSmemT<v3::kStageA>), st>>>(P);
          }
        };
        if (b->n_probe_simple > 0) {
          P.work_query = L.work_query; P.work_slice = L.work_slice; P.n_work = b->n_probe_simple; P.work_counter = b->work_counter.p;
          P.stats = debug ? b->probe_stats.p : nullptr;
          launch(std::true_type{}, cfg_b_simple, b->n_probe_simple);
        }
        if (b->n_probe_generic > 0) {
          P.work_query = L.work_query + b->n_probe_simple; P.work_slice = L.work_slice + b->n_probe_simple;
          P.n_work = b->n_probe_generic; P.work_counter = b->work_counter.p + 1;
          P.stats = debug ? b->probe_stats.p + 16 : nullptr;
          launch(std::false_type{}, cfg_b_generic, b->n_probe_generic);
        }
        NRT_CUDA_TRY(cudaGetLastError());
      }
      if (b->n_stream > 0) {
        v2::StreamLaunch S;
        S.ix = L.ix; S.clauses = L.clauses; S.queries = L.queries;
        S.gbounds = b->gbounds.p; S.n_gran = b->n_gran; S.qtables = b->qtables.p; S.n_slices = L.n_slices; S.top_k = L.top_k;
        S.slice_docs = b->slice_docs;
        S.threshold = b->threshold; S.pruned = b->pruned.p;
        S.mode_stats = nullptr;
        if (debug && !b->use_probe) {
          if (!b->mode_stats.p && (rc_dbg = b->mode_stats.alloc(13))) return rc_dbg;
          NRT_CUDA_TRY(cudaMemsetAsync(b->mode_stats.p, 0, 13 * sizeof(unsigned long long), st));
          S.mode_stats = b->mode_stats.p;
        }
        S.theta = L.theta; S.total_hits = L.total_hits; S.slice_keys = L.slice_keys; S.slice_cnt = L.slice_cnt;
        const int first = n_probe;   // stream items follow the probe items in the work list
        const int n_simple = b->use_probe ? 0 : b->n_work_simple;
        if (n_simple > 0) {
          S.work_query = L.work_query + first; S.work_slice = L.work_slice + first; S.n_work = n_simple;
          v2::posting_stream_kernel<true><<<n_simple, v2::kThreads, sizeof(v2::StreamSmem), st>>>(S);
        }
        if (b->n_stream > n_simple) {
          S.work_query = L.work_query + first + n_simple; S.work_slice = L.work_slice + first + n_simple;
          S.n_work = b->n_stream - n_simple;
          v2::posting_stream_kernel<false><<<S.n_work, v2::kThreads, sizeof(v2::StreamSmem), st>>>(S);
        }
        NRT_CUDA_TRY(cudaGetLastError());
      }
    } else
      bool_window_kernel<uint64_t><<<b->n_work, kThreads, sizeof(BoolSmem<uint64_t>), st>>>(L);
    NRT_CUDA_TRY(cudaGetLastError());
  }
  NRT_CUDA_TRY(cudaEventRecord(ev[1], st));
  if (debug && b->mode_stats.p && !b->use_probe) {
    unsigned long long h[13];
    NRT_CUDA_TRY(cudaMemcpyAsync(h, b->mode_stats.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    NRT_CUDA_TRY(cudaStreamSynchronize(st));
    fprintf(stderr, "[nrtgpu modes] window: %llu items %.0f cyc/item %llu driver postings | window+maxscore: %llu items %.0f %llu | sparse: %llu items %.0f %llu\n",
            h[1], h[1] ? (double)h[0] / h[1] : 0.0, h[6], h[3], h[3] ? (double)h[2] / h[3] : 0.0, h[7], h[5], h[5] ? (double)h[4] / h[5] : 0.0, h[8]);
    if (h[5]) fprintf(stderr, "[nrtgpu modes] sparse items: set-up %.0f cyc, sweep %.0f, flush+output %.0f, %.2f runs/item\n", (double)h[9] / h[5], (double)h[10] / h[5], (double)h[11] / h[5], (double)h[12] / h[5]);
  }
  if (debug && b->probe_stats.p && b->use_probe) {
    unsigned long long h[32];
    NRT_CUDA_TRY(cudaMemcpyAsync(h, b->probe_stats.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    NRT_CUDA_TRY(cudaStreamSynchronize(st));
    for (int k = 0; k < 2; ++k) {
      const unsigned long long* x = h + 16 * k;
      if (x[0]) fprintf(stderr, "[nrtgpu probe %s] longest item %llu cyc; CTA busy: mean %.0f max %llu cyc; warm-up items %llu, %.0f cyc each; per item: flush %.0f cyc (sort %.0f), TMA wait %.0f cyc\n", k == 0 ? "simple" : "generic",
                        x[8], (double)x[9] / std::min<double>((double)x[0], (double)(v3::kCtasA * b->ix->ctx->sm_count)), x[10], x[11], x[11] ? (double)x[12] / x[11] : 0.0, (double)x[13] / x[0], (double)x[15] / x[0], (double)x[14] / x[0]);
      if (x[0]) fprintf(stderr, "[nrtgpu probe %s] %llu items, %.0f cyc/item (set-up %.0f), %.2f runs/item (%.2f staged), %.1f rounds/item, %llu driver postings (%.0f/item), %.2f flushes/item\n",
                        k == 0 ? "simple" : "generic", x[0], (double)x[1] / x[0], (double)x[6] / x[0], (double)x[2] / x[0], (double)x[5] / x[0],
                        (double)x[7] / x[0], x[3], (double)x[3] / x[0], (double)x[4] / x[0]);
    }
  }
  MergeLaunch M;
  M.slice_keys = b->slice_keys.p; M.slice_cnt = b->slice_cnt.p;
  M.n_lists = b->n_lists; M.top_k = b->top_k; M.nq = b->nq; M.doc_base = b->ix->doc_base;
  M.out_docs = b->o_docs(); M.out_scores = b->o_scores(); M.out_counts = b->o_counts();
  M.total_hits = b->total_hits.p; M.pruned = b->pruned.p; M.terminated = b->terminated.p; M.terminate_after = b->ta_scalar;
  M.out_total = b->bound_total; M.out_flags = b->bound_flags;
  M.theta = b->use_probe ? b->theta.p : nullptr;
  M.known_hits = (b->use_probe && !b->ix->live_bits.p) ? b->known_hits.p : nullptr;
  merge_slices_kernel<<<b->nq, kMergeThreads, 0, st>>>(M);
  NRT_CUDA_TRY(cudaGetLastError());
  if (b->sort_kind != NRTGPU_SORT_RELEVANCE) {   // FieldDoc values of the final hits; scores become NaN
    SortValuesLaunch V;
    V.docs = b->o_docs(); V.counts = b->o_counts(); V.nq = b->nq; V.top_k = b->top_k; V.doc_base = b->ix->doc_base; V.kind = b->sort_kind;
    const bool col = b->sort_kind == NRTGPU_SORT_COLUMN;
    V.c64 = col ? b->ix->col64[(size_t)b->sort_column]->p : nullptr; V.c32 = col ? b->ix->col32[(size_t)b->sort_column]->p : nullptr;
    V.has = col ? b->ix->col_has[(size_t)b->sort_column]->p : nullptr; V.missing_value = b->sort_missing_value;
    V.out_values = b->out_sort_values.p; V.out_scores = b->o_scores();
    const int n = b->nq * b->top_k;
    sort_values_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(V);
    NRT_CUDA_TRY(cudaGetLastError());
  }
  NRT_CUDA_TRY(cudaEventRecord(ev[2], st));
  b->runs_recorded++;
  b->ran = true;
  return NRTGPU_OK;
}

static int batch_fetch_impl(nrtgpu_batch* b, void* stream_, int32_t* out_docs, float* out_scores, int32_t* out_counts,
                            int64_t* out_total_hits, uint8_t* out_relation, uint8_t* out_hit_timeout, uint8_t* out_terminated_early) {
  if (!b || !b->ran) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_batch_fetch: batch has not run");
  cudaStream_t st = (cudaStream_t)stream_;
  size_t n = (size_t)b->nq * b->top_k;
  if (out_docs) NRT_CUDA_TRY(cudaMemcpyAsync(out_docs, b->o_docs(), n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (out_scores) NRT_CUDA_TRY(cudaMemcpyAsync(out_scores, b->o_scores(), n * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (out_counts) NRT_CUDA_TRY(cudaMemcpyAsync(out_counts, b->o_counts(), (size_t)b->nq * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (out_total_hits) NRT_CUDA_TRY(cudaMemcpyAsync(out_total_hits, b->total_hits.p, (size_t)b->nq * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  std::vector<int32_t>& pr = b->h_flags;
  pr.resize(3 * (size_t)b->nq);
  NRT_CUDA_TRY(cudaMemcpyAsync(pr.data(), b->pruned.p, (size_t)b->nq * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  NRT_CUDA_TRY(cudaMemcpyAsync(pr.data() + b->nq, b->terminated.p, (size_t)b->nq * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  const bool has_to = b->timed_out.p != nullptr && b->limits_active;
  if (has_to) NRT_CUDA_TRY(cudaMemcpyAsync(pr.data() + 2 * (size_t)b->nq, b->timed_out.p, (size_t)b->nq * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  NRT_CUDA_TRY(cudaStreamSynchronize(st));
  bool any_timeout = false;
  for (int i = 0; i < b->nq; ++i) {
    const bool term = pr[(size_t)b->nq + i] != 0;
    const bool to = has_to && pr[2 * (size_t)b->nq + i] != 0;
    any_timeout |= to;
    // TerminateAfterWrapper.java:85-90: an early-terminated search reports (hits counted, GREATER_THAN_OR_EQUAL_TO)
    if (out_relation) out_relation[i] = (pr[(size_t)i] || term || to) ? 1 : 0;
    if (out_terminated_early) out_terminated_early[i] = term ? 1 : 0;
    if (out_hit_timeout) out_hit_timeout[i] = to ? 1 : 0;
    if (out_total_hits && pr[(size_t)i] && !term && !to && !b->ix->live_bits.p && (size_t)i < b->h_known.size() && (int64_t)b->h_known[(size_t)i] > out_total_hits[i])
      out_total_hits[i] = (int64_t)b->h_known[(size_t)i];   // pruned search: the count is a lower bound; so is the longest list
    if (term && out_total_hits && b->terminate_after_max_recall > 0 && out_total_hits[i] > b->terminate_after_max_recall)
      out_total_hits[i] = b->terminate_after_max_recall;
  }
  // SearchCutoffWrapper.java:164-174: with noPartialResults a timeout is an error (CollectionTimeoutException), else the
  // partial results are returned and hitTimeout is set
  if (any_timeout && b->disallow_partial) NRT_FAIL(NRTGPU_ERR_TIMEOUT, "Search collection exceeded timeout of " + std::to_string(b->timeout_sec) + "s");
  return NRTGPU_OK;
}

int nrtgpu_batch_fetch(nrtgpu_batch* b, void* stream_, int32_t* out_docs, float* out_scores,
                       int32_t* out_counts, int64_t* out_total_hits, uint8_t* out_relation) {
  return batch_fetch_impl(b, stream_, out_docs, out_scores, out_counts, out_total_hits, out_relation, nullptr, nullptr);
}

int nrtgpu_batch_fetch_ex(nrtgpu_batch* b, void* stream_, int32_t* out_docs, float* out_scores, int32_t* out_counts,
                          int64_t* out_total_hits, uint8_t* out_relation, uint8_t* out_hit_timeout, uint8_t* out_terminated_early) {
  return batch_fetch_impl(b, stream_, out_docs, out_scores, out_counts, out_total_hits, out_relation, out_hit_timeout, out_terminated_early);
}

// aggregation results of the last run -> caller buffers
static int batch_fetch_aggs(nrtgpu_batch* b, cudaStream_t st, const nrtgpu_aggregation_result* out) {
  const int nq = b->nq;
  for (size_t i = 0; i < b->aggs.size(); ++i) {
    const nrtgpu_aggregation& a = b->aggs[i];
    const nrtgpu_aggregation_result& r = out[i];
    if (a.kind == NRTGPU_AGG_TERMS) {
      int rc;
      const size_t n = (size_t)nq * a.size;
      if ((rc = b->agg_keys.alloc(n)) || (rc = b->agg_cnts.alloc(n)) || (rc = b->agg_n.alloc((size_t)nq)) || (rc = b->agg_tot.alloc((size_t)nq)) ||
          (rc = b->agg_other.alloc((size_t)nq))) return rc;
      AggTermsLaunch T;
      T.counts = b->agg_counts[i].p; T.n_buckets = b->ix->col_n_distinct[(size_t)a.column]; T.nq = nq; T.size = a.size; T.order_desc = a.order_desc != 0;
      T.distinct = b->ix->col_distinct[(size_t)a.column]->p;
      T.out_keys = b->agg_keys.p; T.out_counts = b->agg_cnts.p; T.out_n = b->agg_n.p; T.out_total_buckets = b->agg_tot.p; T.out_other = b->agg_other.p;
      agg_terms_topk_kernel<<<nq, 256, 0, st>>>(T);
      NRT_CUDA_TRY(cudaGetLastError());
      if (r.bucket_keys) NRT_CUDA_TRY(cudaMemcpyAsync(r.bucket_keys, b->agg_keys.p, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      if (r.bucket_counts) NRT_CUDA_TRY(cudaMemcpyAsync(r.bucket_counts, b->agg_cnts.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      if (r.n_buckets) NRT_CUDA_TRY(cudaMemcpyAsync(r.n_buckets, b->agg_n.p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      if (r.total_buckets) NRT_CUDA_TRY(cudaMemcpyAsync(r.total_buckets, b->agg_tot.p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      if (r.other_counts) NRT_CUDA_TRY(cudaMemcpyAsync(r.other_counts, b->agg_other.p, (size_t)nq * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      NRT_CUDA_TRY(cudaStreamSynchronize(st));   // the scratch is reused by the next terms aggregation
    } else if (r.values) {
      std::vector<unsigned long long> h((size_t)nq);
      NRT_CUDA_TRY(cudaMemcpyAsync(h.data(), b->agg_dvals[i].p, (size_t)nq * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
      NRT_CUDA_TRY(cudaStreamSynchronize(st));
      for (int q = 0; q < nq; ++q) {
        double v;
        if (a.kind == NRTGPU_AGG_SUM) std::memcpy(&v, &h[(size_t)q], sizeof(v));
        else if (a.kind == NRTGPU_AGG_MAX) v = h[(size_t)q] == 0ull ? -DBL_MAX : ordered_to_double(h[(size_t)q]);              // MaxCollectorManager.UNSET_VALUE
        else v = h[(size_t)q] == 0xffffffffffffffffull ? DBL_MAX : ordered_to_double(h[(size_t)q]);                          // MinCollectorManager.UNSET_VALUE
        r.values[q] = v;
      }
    }
  }
  return NRTGPU_OK;
}

// deadline / terminateAfter of the batch (SearchCutoffWrapper / TerminateAfterWrapper semantics, see include/nrtgpu.h)
static int batch_set_limits(nrtgpu_batch* b, const nrtgpu_search_limits* lim, cudaStream_t st) {
  b->limits_active = false; b->disallow_partial = false; b->timeout_sec = 0.0; b->terminate_after_max_recall = 0;
  b->deadline_ns = 0; b->ta_scalar = 0;
  if (!lim) return NRTGPU_OK;
  if (lim->timeout_sec < 0.0 || lim->terminate_after < 0) NRT_FAIL(NRTGPU_ERR_INVALID, "timeout_sec / terminate_after must be >= 0");
  int rc;
  if (lim->timeout_sec > 0.0) {
    b->limits_active = true; b->timeout_sec = lim->timeout_sec; b->disallow_partial = lim->disallow_partial_results != 0;
    const double left = lim->timeout_sec - lim->elapsed_sec;   // the timer started when the request's first collector was created
    b->deadline_ns = left <= 0.0 ? -1 : std::max<long long>(1, (long long)(left * 1e9));
    if ((rc = b->timed_out.alloc((size_t)b->nq))) return rc;
    if ((rc = b->clock0.alloc(1))) return rc;
  }
  if (lim->terminate_after > 0) {
    b->ta_scalar = lim->terminate_after;
    b->terminate_after_max_recall = lim->terminate_after_max_recall_count > lim->terminate_after ? lim->terminate_after_max_recall_count : lim->terminate_after;
  }
  (void)st;
  return NRTGPU_OK;
}

int nrtgpu_batch_set_limits(nrtgpu_batch* b, const nrtgpu_search_limits* limits) {
  if (!b) NRT_FAIL(NRTGPU_ERR_INVALID, "NULL batch");
  return batch_set_limits(b, limits, (cudaStream_t)0);
}

int nrtgpu_batch_device_results(nrtgpu_batch* b, int32_t** d_docs, float** d_scores, int32_t** d_counts) {
  if (!b) NRT_FAIL(NRTGPU_ERR_INVALID, "NULL batch");
  if (d_docs) *d_docs = b->o_docs();
  if (d_scores) *d_scores = b->o_scores();
  if (d_counts) *d_counts = b->o_counts();
  return NRTGPU_OK;
}

int nrtgpu_batch_bind_output(nrtgpu_batch* b, int32_t* d_docs, float* d_scores, int32_t* d_counts) {
  if (!b) NRT_FAIL(NRTGPU_ERR_INVALID, "NULL batch");
  b->bound_docs = d_docs; b->bound_scores = d_scores; b->bound_counts = d_counts;
  b->bound_total = nullptr; b->bound_flags = nullptr;
  return NRTGPU_OK;
}

// Packed per-shard result record (one all-gather carries everything TopDocs.merge needs), int32 words:
//   docs [nq*top_k] | scores [nq*top_k] (float bits) | counts [nq] | flags [nq] | (pad to 8 bytes) | totalHits [nq] int64
int64_t nrtgpu_packed_words(int32_t nq, int32_t top_k) {
  int64_t w = (int64_t)nq * top_k * 2 + 2ll * nq;
  w = (w + 1) & ~1ll;
  return w + 2ll * nq;
}

int nrtgpu_batch_bind_packed(nrtgpu_batch* b, int32_t* d_record) {
  if (!b) NRT_FAIL(NRTGPU_ERR_INVALID, "NULL batch");
  if (!d_record) { b->bound_docs = nullptr; b->bound_scores = nullptr; b->bound_counts = nullptr; b->bound_total = nullptr; b->bound_flags = nullptr; return NRTGPU_OK; }
  if (((uintptr_t)d_record & 7u) != 0) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_batch_bind_packed: record must be 8-byte aligned");
  const int64_t n = (int64_t)b->nq * b->top_k;
  b->bound_docs = d_record; b->bound_scores = (float*)(d_record + n); b->bound_counts = d_record + 2 * n;
  b->bound_flags = d_record + 2 * n + b->nq;
  int64_t w = 2 * n + 2ll * b->nq; w = (w + 1) & ~1ll;
  b->bound_total = (long long*)(d_record + w);
  return NRTGPU_OK;
}

int nrtgpu_merge_topk_packed(nrtgpu_ctx* ctx, int32_t n_lists, int32_t nq, int32_t top_k, const int32_t* d_records,
                             int32_t* d_out_record, void* stream) {
  if (!ctx || !d_records || !d_out_record || n_lists <= 0 || nq <= 0 || top_k <= 0 || top_k > kMaxTopK)
    NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_merge_topk_packed: bad argument");
  NRT_CUDA_TRY(cudaSetDevice(ctx->device));
  const int64_t n = (int64_t)nq * top_k, words = nrtgpu_packed_words(nq, top_k);
  int64_t w = 2 * n + 2ll * nq; w = (w + 1) & ~1ll;
  MergePairsLaunch M;
  M.docs = d_records; M.scores = (const float*)(d_records + n); M.counts = d_records + 2 * n;
  M.stride_hits = words; M.stride_counts = words;
  M.n_lists = n_lists; M.top_k = top_k; M.nq = nq;
  M.out_docs = d_out_record; M.out_scores = (float*)(d_out_record + n); M.out_counts = d_out_record + 2 * n;
  M.flags = d_records + 2 * n + nq; M.totals = (const long long*)(d_records + w); M.stride_flags = words; M.stride_totals = words / 2;
  M.out_flags = d_out_record + 2 * n + nq; M.out_total = (long long*)(d_out_record + w);
  merge_pairs_kernel<<<nq, kMergeThreads, 0, (cudaStream_t)stream>>>(M);
  NRT_CUDA_TRY(cudaGetLastError());
  return NRTGPU_OK;
}

int nrtgpu_batch_reset_timing(nrtgpu_batch* b) {
  if (!b) NRT_FAIL(NRTGPU_ERR_INVALID, "NULL batch");
  b->runs_recorded = 0;
  return NRTGPU_OK;
}

int nrtgpu_batch_stats(const nrtgpu_batch* b, int64_t* alg_postings, int32_t* launches_per_run, int64_t* work_items) {
  if (!b) NRT_FAIL(NRTGPU_ERR_INVALID, "NULL batch");
  if (alg_postings) *alg_postings = b->alg_postings;
  if (launches_per_run) {
    int n = 1;   // slice merge
    if (b->wide_slots) n += b->n_work > 0 ? 1 : 0;
    else {
      n += (b->n_probe_simple > 0 ? 1 : 0) + (b->n_probe_generic > 0 ? 1 : 0);
      const int n_simple = b->use_probe ? 0 : b->n_work_simple;
      n += (n_simple > 0 ? 1 : 0) + (b->n_stream > n_simple ? 1 : 0);
    }
    *launches_per_run = n;
  }
  if (work_items) *work_items = b->n_work;
  return NRTGPU_OK;
}

int nrtgpu_batch_stage_ms(nrtgpu_batch* b, int32_t stage, float* ms) {
  if (!b || !ms || stage < 0 || stage > 1 || !b->ran || b->runs_recorded < 1) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_batch_stage_ms: bad argument");
  int n = std::min(b->runs_recorded, (int)nrtgpu_batch::kEvRing);
  double sum = 0.0;
  for (int i = 0; i < n; ++i) {
    float t = 0.f;
    NRT_CUDA_TRY(cudaEventElapsedTime(&t, b->ev[i][stage], b->ev[i][stage + 1]));
    sum += t;
  }
  *ms = (float)(sum / n);
  return NRTGPU_OK;
}

int nrtgpu_batch_free(nrtgpu_batch* b) {
  if (b) { cudaSetDevice(b->ix->ctx->device); delete b; }
  return NRTGPU_OK;
}

// one-shot search: compile + upload the batch into a pooled workspace, run; results either copied to HOST buffers
// (d_record == NULL) or left in a packed DEVICE record (the multi-GPU path: the caller all-gathers it on `stream`)
static int search_bool_impl(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                            const nrtgpu_query* queries, int32_t nq, int32_t top_k, int32_t total_hits_threshold, int32_t flags,
                            const nrtgpu_search_limits* limits, void* stream, int32_t* d_record, int32_t* out_docs, float* out_scores,
                            int32_t* out_counts, int64_t* out_total_hits, uint8_t* out_relation, uint8_t* out_hit_timeout,
                            uint8_t* out_terminated_early, const nrtgpu_sort* sort = nullptr, int64_t* out_sort_values = nullptr,
                            const nrtgpu_aggregation* aggs = nullptr, int32_t n_aggs = 0, const nrtgpu_aggregation_result* agg_out = nullptr) {
  if (!ix) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_search_bool: NULL index");
  // take a cached workspace (device buffers survive between calls: no cudaMalloc on the request path)
  nrtgpu_batch* b = nullptr;
  {
    std::lock_guard<std::mutex> g(ix->ws_mu);
    if (!ix->ws_free.empty()) { b = ix->ws_free.back(); ix->ws_free.pop_back(); }
  }
  if (!b) b = new nrtgpu_batch;
  b->bound_docs = nullptr; b->bound_scores = nullptr; b->bound_counts = nullptr; b->bound_total = nullptr; b->bound_flags = nullptr;
  int rc = batch_build(b, ix, clauses, n_clauses, queries, nq, top_k, total_hits_threshold, flags, (cudaStream_t)stream, sort, aggs, n_aggs);
  if (!rc) rc = batch_set_limits(b, limits, (cudaStream_t)stream);
  if (!rc && d_record) rc = nrtgpu_batch_bind_packed(b, d_record);
  if (!rc) rc = nrtgpu_batch_run(b, stream);
  if (!rc) {
    if (d_record) { cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream); if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); rc = NRTGPU_ERR_CUDA; } }
    else {
      if (out_sort_values && b->sort_kind != NRTGPU_SORT_RELEVANCE) {
        cudaError_t e = cudaMemcpyAsync(out_sort_values, b->out_sort_values.p, (size_t)nq * top_k * sizeof(int64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); rc = NRTGPU_ERR_CUDA; }
      }
      if (!rc) rc = batch_fetch_impl(b, stream, out_docs, out_scores, out_counts, out_total_hits, out_relation, out_hit_timeout, out_terminated_early);
      if (!rc && n_aggs > 0 && agg_out) rc = batch_fetch_aggs(b, (cudaStream_t)stream, agg_out);
    }
  }
  b->bound_docs = nullptr; b->bound_scores = nullptr; b->bound_counts = nullptr; b->bound_total = nullptr; b->bound_flags = nullptr;
  {
    std::lock_guard<std::mutex> g(ix->ws_mu);
    ix->ws_free.push_back(b);
  }
  return rc;
}

int nrtgpu_search_bool(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                       const nrtgpu_query* queries, int32_t nq, int32_t top_k,
                       int32_t total_hits_threshold, int32_t flags, void* stream, int32_t* out_docs,
                       float* out_scores, int32_t* out_counts, int64_t* out_total_hits,
                       uint8_t* out_relation) {
  return search_bool_impl(ix, clauses, n_clauses, queries, nq, top_k, total_hits_threshold, flags, nullptr, stream, nullptr,
                          out_docs, out_scores, out_counts, out_total_hits, out_relation, nullptr, nullptr);
}

int nrtgpu_search_bool_ex(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                          const nrtgpu_query* queries, int32_t nq, int32_t top_k, int32_t total_hits_threshold, int32_t flags,
                          const nrtgpu_search_limits* limits, void* stream, int32_t* out_docs, float* out_scores,
                          int32_t* out_counts, int64_t* out_total_hits, uint8_t* out_relation, uint8_t* out_hit_timeout,
                          uint8_t* out_terminated_early) {
  return search_bool_impl(ix, clauses, n_clauses, queries, nq, top_k, total_hits_threshold, flags, limits, stream, nullptr,
                          out_docs, out_scores, out_counts, out_total_hits, out_relation, out_hit_timeout, out_terminated_early);
}

int nrtgpu_search_sorted(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                         const nrtgpu_query* queries, int32_t nq, int32_t top_k, int32_t flags,
                         const nrtgpu_sort* sort, const nrtgpu_search_limits* limits, void* stream,
                         int32_t* out_docs, int64_t* out_sort_values, int32_t* out_counts,
                         int64_t* out_total_hits, uint8_t* out_relation, uint8_t* out_hit_timeout,
                         uint8_t* out_terminated_early) {
  if (!sort) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_search_sorted: NULL sort");
  return search_bool_impl(ix, clauses, n_clauses, queries, nq, top_k, INT32_MAX, flags, limits, stream, nullptr,
                          out_docs, nullptr, out_counts, out_total_hits, out_relation, out_hit_timeout, out_terminated_early,
                          sort, out_sort_values);
}

int nrtgpu_search_bool_aggs(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                            const nrtgpu_query* queries, int32_t nq, int32_t top_k, int32_t flags,
                            const nrtgpu_aggregation* aggs, int32_t n_aggs, const nrtgpu_aggregation_result* results,
                            void* stream, int32_t* out_docs, float* out_scores, int32_t* out_counts,
                            int64_t* out_total_hits) {
  if (n_aggs <= 0 || !aggs || !results) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_search_bool_aggs: no aggregations");
  return search_bool_impl(ix, clauses, n_clauses, queries, nq, top_k, INT32_MAX, flags, nullptr, stream, nullptr, out_docs, out_scores,
                          out_counts, out_total_hits, nullptr, nullptr, nullptr, nullptr, nullptr, aggs, n_aggs, results);
}

int nrtgpu_score_docs(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                      const nrtgpu_query* queries, int32_t nq, int32_t n_hits, const int32_t* docs,
                      const int32_t* counts, void* stream, uint8_t* out_matches, float* out_scores) {
  if (!ix || !docs || !out_matches || !out_scores || n_hits <= 0) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_score_docs: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  nrtgpu_batch* b = nullptr;
  { std::lock_guard<std::mutex> g(ix->ws_mu); if (!ix->ws_free.empty()) { b = ix->ws_free.back(); ix->ws_free.pop_back(); } }
  if (!b) b = new nrtgpu_batch;
  int rc = batch_build(b, ix, clauses, n_clauses, queries, nq, 1, INT32_MAX, 0, st);
  const size_t n = (size_t)nq * n_hits;
  if (!rc) rc = b->sd_docs.upload_async(docs, n, st);
  if (!rc && counts) rc = b->sd_counts.upload_async(counts, (size_t)nq, st);
  if (!rc) rc = b->sd_match.alloc(n);
  if (!rc) rc = b->sd_scores.alloc(n);
  if (!rc) {
    ScoreDocsLaunch S; S.ix = ix->view(); S.clauses = b->clauses.p; S.queries = b->queries.p; S.nq = nq; S.n_hits = n_hits;
    S.docs = b->sd_docs.p; S.counts = counts ? b->sd_counts.p : nullptr; S.out_matches = b->sd_match.p; S.out_scores = b->sd_scores.p;
    score_docs_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(S);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_matches, b->sd_match.p, n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_scores, b->sd_scores.p, n * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); rc = NRTGPU_ERR_CUDA; }
  }
  { std::lock_guard<std::mutex> g(ix->ws_mu); ix->ws_free.push_back(b); }
  return rc;
}

int nrtgpu_rescore_query(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                         const nrtgpu_query* queries, int32_t nq, int32_t n_hits, const int32_t* counts,
                         int32_t window, double query_weight, double rescore_weight, void* stream,
                         int32_t* docs, float* scores, int32_t* out_counts) {
  if (!ix || !docs || !scores || n_hits <= 0 || window <= 0) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_rescore_query: bad argument");
  if (n_hits > kHybCap) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "nrtgpu_rescore_query: more than 4096 hits per query");
  cudaStream_t st = (cudaStream_t)stream;
  nrtgpu_batch* b = nullptr;
  { std::lock_guard<std::mutex> g(ix->ws_mu); if (!ix->ws_free.empty()) { b = ix->ws_free.back(); ix->ws_free.pop_back(); } }
  if (!b) b = new nrtgpu_batch;
  int rc = batch_build(b, ix, clauses, n_clauses, queries, nq, 1, INT32_MAX, 0, st);
  const size_t n = (size_t)nq * n_hits;
  // Lucene QueryRescorer.rescore(searcher, hits, topN = windowSize): EVERY first-pass hit is combined and the list
  // re-sorted (score desc, doc asc); then the first topN are kept
  std::vector<int32_t> wc((size_t)nq);
  for (int q = 0; q < nq; ++q) {
    wc[(size_t)q] = counts ? counts[q] : n_hits;
    if (wc[(size_t)q] < 0 || wc[(size_t)q] > n_hits) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_rescore_query: counts out of range");
  }
  if (!rc) rc = b->sd_docs.upload_async(docs, n, st);
  if (!rc) rc = b->sd_first.upload_async(scores, n, st);
  if (!rc) rc = b->sd_counts.upload_async(wc.data(), (size_t)nq, st);
  if (!rc) rc = b->sd_match.alloc(n);
  if (!rc) rc = b->sd_scores.alloc(n);
  if (!rc) {
    ScoreDocsLaunch S; S.ix = ix->view(); S.clauses = b->clauses.p; S.queries = b->queries.p; S.nq = nq; S.n_hits = n_hits;
    S.docs = b->sd_docs.p; S.counts = b->sd_counts.p; S.out_matches = b->sd_match.p; S.out_scores = b->sd_scores.p;
    score_docs_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(S);
    RescoreLaunch P;
    P.nq = nq; P.n_hits = n_hits; P.counts = b->sd_counts.p; P.docs = b->sd_docs.p; P.scores = b->sd_first.p;
    P.second_matches = b->sd_match.p; P.second_scores = b->sd_scores.p; P.query_weight = query_weight; P.rescore_weight = rescore_weight;
    rescore_combine_kernel<<<nq, kHybThreads, 0, st>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(docs, b->sd_docs.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(scores, b->sd_first.p, n * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); rc = NRTGPU_ERR_CUDA; }
    if (!rc && out_counts) for (int q = 0; q < nq; ++q) out_counts[q] = std::min(wc[(size_t)q], window);
  }
  { std::lock_guard<std::mutex> g(ix->ws_mu); ix->ws_free.push_back(b); }
  return rc;
}

int nrtgpu_fetch_columns(nrtgpu_index* ix, const int32_t* col_ids, int32_t n_cols, const int32_t* docs, int32_t n,
                         void* stream, int64_t* out_values, uint8_t* out_has) {
  if (!ix || !col_ids || !docs || !out_values || !out_has || n_cols <= 0 || n <= 0) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_fetch_columns: bad argument");
  for (int i = 0; i < n_cols; ++i) if (col_ids[i] < 0 || col_ids[i] >= ix->n_columns) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_fetch_columns: column out of range");
  for (int i = 0; i < n_cols; ++i) if (ix->col_multi[(size_t)col_ids[i]]) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "nrtgpu_fetch_columns: multi-valued column");
  NRT_CUDA_TRY(cudaSetDevice(ix->ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  std::lock_guard<std::mutex> g(ix->fetch_mu);
  int rc;
  const size_t total = (size_t)n_cols * (size_t)n;
  if ((rc = ix->f_cols.upload_async(col_ids, (size_t)n_cols, st)) || (rc = ix->f_docs.upload_async(docs, (size_t)n, st)) ||
      (rc = ix->f_vals.alloc(total)) || (rc = ix->f_has.alloc(total))) return rc;
  FetchLaunch F; F.ix = ix->view(); F.col_ids = ix->f_cols.p; F.n_cols = n_cols; F.docs = ix->f_docs.p; F.n = n;
  F.out_values = ix->f_vals.p; F.out_has = ix->f_has.p;
  fetch_columns_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(F);
  NRT_CUDA_TRY(cudaGetLastError());
  NRT_CUDA_TRY(cudaMemcpyAsync(out_values, ix->f_vals.p, total * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  NRT_CUDA_TRY(cudaMemcpyAsync(out_has, ix->f_has.p, total, cudaMemcpyDeviceToHost, st));
  NRT_CUDA_TRY(cudaStreamSynchronize(st));
  return NRTGPU_OK;
}

int nrtgpu_search_bool_packed(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                              const nrtgpu_query* queries, int32_t nq, int32_t top_k, int32_t total_hits_threshold, int32_t flags,
                              const nrtgpu_search_limits* limits, void* stream, int32_t* d_record) {
  if (!d_record) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_search_bool_packed: NULL record");
  return search_bool_impl(ix, clauses, n_clauses, queries, nq, top_k, total_hits_threshold, flags, limits, stream, d_record,
                          nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int nrtgpu_merge_topk_device(nrtgpu_ctx* ctx, int32_t n_lists, int32_t nq, int32_t top_k,
                             const int32_t* d_docs, const float* d_scores, const int32_t* d_counts,
                             int32_t* d_out_docs, float* d_out_scores, int32_t* d_out_counts, void* stream) {
  if (!ctx || n_lists <= 0 || nq <= 0 || top_k <= 0 || top_k > kMaxTopK) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_merge_topk_device: bad argument");
  NRT_CUDA_TRY(cudaSetDevice(ctx->device));
  MergePairsLaunch M;
  M.docs = d_docs; M.scores = d_scores; M.counts = d_counts; M.n_lists = n_lists; M.top_k = top_k; M.nq = nq;
  M.stride_hits = (int64_t)nq * top_k; M.stride_counts = nq;
  M.out_docs = d_out_docs; M.out_scores = d_out_scores; M.out_counts = d_out_counts;
  M.totals = nullptr; M.flags = nullptr; M.stride_totals = 0; M.stride_flags = 0; M.out_total = nullptr; M.out_flags = nullptr;
  merge_pairs_kernel<<<nq, kMergeThreads, 0, (cudaStream_t)stream>>>(M);
  NRT_CUDA_TRY(cudaGetLastError());
  return NRTGPU_OK;
}

int nrtgpu_search_knn(nrtgpu_index* ix, const float* queries, int32_t nq, int32_t k, const float* boosts,
                      const uint8_t* filter, void* stream, int32_t* out_docs, float* out_scores,
                      int32_t* out_counts) {
  if (!ix || !queries || !out_docs || !out_scores || !out_counts) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_search_knn: NULL argument");
  if (ix->vec_dims <= 0) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_search_knn: index has no vector field");
  if (k <= 0 || k > kMaxTopK) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_search_knn: k out of range");
  NRT_CUDA_TRY(cudaSetDevice(ix->ctx->device));
  const bool tcp = ix->vec_tc && !(getenv("NRTGPU_KNN_SIMT") != nullptr);
  std::lock_guard<std::mutex> g(ix->knn_mu);
  return knn_search_host(ix->vectors.p, ix->vec_norm2.p, ix->vec_docs.p, ix->vec_count, ix->vec_dims, ix->vec_sim | (ix->vec_is_byte ? kKnnByteFlag : 0),
                         ix->doc_base, ix->n_docs, queries, nq, k, boosts, filter, (cudaStream_t)stream, out_docs,
                         out_scores, out_counts, tcp ? ix->vec_bf16.p : nullptr, tcp ? &ix->vec_tmap : nullptr, nullptr, ix->vec_ab.p,
                         &ix->knn_scratch, ix->live_bits.p, ix->vec_dmax, &ix->knn_last_uncertified, tcp ? &ix->vec_tmap128 : nullptr);
}

int nrtgpu_search_knn_timed(nrtgpu_index* ix, const float* queries, int32_t nq, int32_t k, void* stream, int32_t* out_docs,
                            float* out_scores, int32_t* out_counts, float* stage_ms /*[3]: gemm, select, rescore*/) {
  if (!ix || !queries || !out_docs || !out_scores || !out_counts || !stage_ms) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_search_knn_timed: NULL argument");
  if (ix->vec_dims <= 0) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_search_knn_timed: index has no vector field");
  if (k <= 0 || k > kMaxTopK / 4) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_search_knn_timed: k out of range");
  NRT_CUDA_TRY(cudaSetDevice(ix->ctx->device));
  const bool tcp = ix->vec_tc && !(getenv("NRTGPU_KNN_SIMT") != nullptr);
  std::lock_guard<std::mutex> g(ix->knn_mu);
  return knn_search_host(ix->vectors.p, ix->vec_norm2.p, ix->vec_docs.p, ix->vec_count, ix->vec_dims, ix->vec_sim | (ix->vec_is_byte ? kKnnByteFlag : 0),
                         ix->doc_base, ix->n_docs, queries, nq, k, nullptr, nullptr, (cudaStream_t)stream, out_docs,
                         out_scores, out_counts, tcp ? ix->vec_bf16.p : nullptr, tcp ? &ix->vec_tmap : nullptr, stage_ms, ix->vec_ab.p,
                         &ix->knn_scratch, ix->live_bits.p, ix->vec_dmax, &ix->knn_last_uncertified, tcp ? &ix->vec_tmap128 : nullptr);
}

int32_t nrtgpu_knn_last_uncertified(const nrtgpu_index* ix) { return ix ? ix->knn_last_uncertified : 0; }

// weighted RRF (mode 0) or score-order (MAX / SUM / AVG) blend of R retrievers' lists; HOST buffers, pooled device scratch
static int blend_impl(nrtgpu_ctx* ctx, int32_t mode, int32_t R, int32_t nq, int32_t top_in, const int32_t* docs, const float* scores,
                      const int32_t* counts, const float* boosts, int32_t rank_constant, int32_t top_out, int32_t* out_docs,
                      float* out_scores, int32_t* out_counts, int32_t* out_total) {
  if (!ctx || !docs || !counts || !boosts || !out_docs || !out_scores || !out_counts || !out_total || (mode != 0 && !scores))
    NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_blend: NULL argument");
  if (R <= 0 || nq <= 0 || top_in <= 0 || top_out <= 0) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_blend: sizes must be > 0");
  if ((int64_t)R * top_in > kHybCap || top_in > 65535 || R > 65535) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "nrtgpu_blend: more than 4096 hits per query");
  for (int64_t i = 0; i < (int64_t)R * nq; ++i)
    if (counts[i] < 0 || counts[i] > top_in) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_blend: counts[] outside [0, top_in]");
  // WeightedRrfBlenderOperation.java:47-49: rankConstant <= 0 selects DEFAULT_K = 60
  const int k = rank_constant > 0 ? rank_constant : 60;
  NRT_CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t nd = (size_t)R * nq * top_in, nc = (size_t)R * nq, no = (size_t)nq * top_out;
  std::lock_guard<std::mutex> g(ctx->hyb_mu);
  int rc;
  // one pooled allocation, carved up (all parts 4-byte types)
  const size_t words = nd * 2 + nc + (size_t)R + no * 2 + (size_t)nq * 2;
  if ((rc = ctx->hyb_scratch.alloc(words))) return rc;
  int32_t* p = ctx->hyb_scratch.p;
  int32_t* d_docs = p; p += nd;
  float* d_scores = (float*)p; p += nd;
  int32_t* d_counts = p; p += nc;
  float* d_boosts = (float*)p; p += R;
  int32_t* d_od = p; p += no;
  float* d_os = (float*)p; p += no;
  int32_t* d_oc = p; p += nq;
  int32_t* d_ot = p;
  cudaStream_t st = 0;
  NRT_CUDA_TRY(cudaMemcpyAsync(d_docs, docs, nd * 4, cudaMemcpyHostToDevice, st));
  if (mode != 0) NRT_CUDA_TRY(cudaMemcpyAsync(d_scores, scores, nd * 4, cudaMemcpyHostToDevice, st));
  NRT_CUDA_TRY(cudaMemcpyAsync(d_counts, counts, nc * 4, cudaMemcpyHostToDevice, st));
  NRT_CUDA_TRY(cudaMemcpyAsync(d_boosts, boosts, (size_t)R * 4, cudaMemcpyHostToDevice, st));
  RrfLaunch P;
  P.docs = d_docs; P.counts = d_counts; P.boosts = d_boosts; P.scores = mode != 0 ? d_scores : nullptr; P.mode = mode;
  P.R = R; P.nq = nq; P.top_in = top_in; P.rank_constant = k; P.top_out = top_out;
  P.out_docs = d_od; P.out_scores = d_os; P.out_counts = d_oc; P.out_total = d_ot;
  rrf_blend_kernel<<<nq, kHybThreads, 0, st>>>(P);
  NRT_CUDA_TRY(cudaGetLastError());
  NRT_CUDA_TRY(cudaMemcpyAsync(out_docs, d_od, no * 4, cudaMemcpyDeviceToHost, st));
  NRT_CUDA_TRY(cudaMemcpyAsync(out_scores, d_os, no * 4, cudaMemcpyDeviceToHost, st));
  NRT_CUDA_TRY(cudaMemcpyAsync(out_counts, d_oc, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  NRT_CUDA_TRY(cudaMemcpyAsync(out_total, d_ot, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  NRT_CUDA_TRY(cudaStreamSynchronize(st));
  return NRTGPU_OK;
}

int nrtgpu_blend_rrf(nrtgpu_ctx* ctx, int32_t R, int32_t nq, int32_t top_in, const int32_t* docs, const int32_t* counts,
                     const float* boosts, int32_t rank_constant, int32_t top_out, int32_t* out_docs, float* out_scores,
                     int32_t* out_counts, int32_t* out_total) {
  return blend_impl(ctx, 0, R, nq, top_in, docs, nullptr, counts, boosts, rank_constant, top_out, out_docs, out_scores, out_counts, out_total);
}

int nrtgpu_blend_scores(nrtgpu_ctx* ctx, int32_t score_mode, int32_t R, int32_t nq, int32_t top_in, const int32_t* docs,
                        const float* scores, const int32_t* counts, const float* boosts, int32_t top_out, int32_t* out_docs,
                        float* out_scores, int32_t* out_counts, int32_t* out_total) {
  if (score_mode < NRTGPU_BLEND_MAX || score_mode > NRTGPU_BLEND_AVG) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_blend_scores: bad score mode");
  return blend_impl(ctx, score_mode, R, nq, top_in, docs, scores, counts, boosts, 0, top_out, out_docs, out_scores, out_counts, out_total);
}

int nrtgpu_rescore_combine(nrtgpu_ctx* ctx, int32_t nq, int32_t n_hits, const int32_t* counts, int32_t* docs, float* scores,
                           const uint8_t* second_matches, const float* second_scores, double query_weight,
                           double rescore_weight) {
  if (!ctx || !docs || !scores || !second_matches || !second_scores) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_rescore_combine: NULL argument");
  if (nq <= 0 || n_hits <= 0) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_rescore_combine: sizes must be > 0");
  if (n_hits > kHybCap) NRT_FAIL(NRTGPU_ERR_UNSUPPORTED, "nrtgpu_rescore_combine: more than 4096 hits per query");
  if (counts) for (int q = 0; q < nq; ++q) if (counts[q] < 0 || counts[q] > n_hits) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_rescore_combine: counts[] outside [0, n_hits]");
  NRT_CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t n = (size_t)nq * n_hits;
  std::lock_guard<std::mutex> g(ctx->hyb_mu);
  int rc;
  if ((rc = ctx->hyb_scratch.alloc(n * 3 + (n + 3) / 4 + (size_t)nq))) return rc;
  int32_t* p = ctx->hyb_scratch.p;
  int32_t* d_docs = p; p += n;
  float* d_scores = (float*)p; p += n;
  float* d_second = (float*)p; p += n;
  int32_t* d_counts = p; p += nq;
  uint8_t* d_match = (uint8_t*)p;
  cudaStream_t st = 0;
  NRT_CUDA_TRY(cudaMemcpyAsync(d_docs, docs, n * 4, cudaMemcpyHostToDevice, st));
  NRT_CUDA_TRY(cudaMemcpyAsync(d_scores, scores, n * 4, cudaMemcpyHostToDevice, st));
  NRT_CUDA_TRY(cudaMemcpyAsync(d_match, second_matches, n, cudaMemcpyHostToDevice, st));
  NRT_CUDA_TRY(cudaMemcpyAsync(d_second, second_scores, n * 4, cudaMemcpyHostToDevice, st));
  if (counts) NRT_CUDA_TRY(cudaMemcpyAsync(d_counts, counts, (size_t)nq * 4, cudaMemcpyHostToDevice, st));
  RescoreLaunch P;
  P.nq = nq; P.n_hits = n_hits; P.counts = counts ? d_counts : nullptr;
  P.docs = d_docs; P.scores = d_scores; P.second_matches = d_match; P.second_scores = d_second;
  P.query_weight = query_weight; P.rescore_weight = rescore_weight;
  rescore_combine_kernel<<<nq, kHybThreads, 0, st>>>(P);
  NRT_CUDA_TRY(cudaGetLastError());
  NRT_CUDA_TRY(cudaMemcpyAsync(docs, d_docs, n * 4, cudaMemcpyDeviceToHost, st));
  NRT_CUDA_TRY(cudaMemcpyAsync(scores, d_scores, n * 4, cudaMemcpyDeviceToHost, st));
  NRT_CUDA_TRY(cudaStreamSynchronize(st));
  return NRTGPU_OK;
}

// ---- searcher over several leaf images of one shard (NRT: a new reader version adds images for the NEW leaves only)
struct nrtgpu_searcher {
  nrtgpu_ctx* ctx = nullptr;
  std::vector<nrtgpu_index*> leaves;
  std::mutex mu;
  DevBuf<int32_t> records, merged;   // [n_leaves][words], [words]
  std::vector<int32_t> host;
};

int nrtgpu_searcher_create(nrtgpu_ctx* ctx, nrtgpu_index* const* leaves, int32_t n_leaves, nrtgpu_searcher** out) {
  if (!ctx || !leaves || n_leaves <= 0 || !out) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_searcher_create: bad argument");
  std::unique_ptr<nrtgpu_searcher> s(new nrtgpu_searcher);
  s->ctx = ctx;
  for (int i = 0; i < n_leaves; ++i) {
    if (!leaves[i] || leaves[i]->ctx != ctx) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_searcher_create: leaf of another context");
    s->leaves.push_back(leaves[i]);
  }
  *out = s.release();
  return NRTGPU_OK;
}

int nrtgpu_searcher_close(nrtgpu_searcher* s) { delete s; return NRTGPU_OK; }

int nrtgpu_searcher_search_bool(nrtgpu_searcher* s, const nrtgpu_clause* clauses, int32_t n_clauses,
                                const nrtgpu_query* queries, int32_t nq, int32_t top_k, int32_t total_hits_threshold,
                                int32_t flags, const nrtgpu_search_limits* limits, void* stream, int32_t* out_docs,
                                float* out_scores, int32_t* out_counts, int64_t* out_total_hits, uint8_t* out_relation) {
  if (!s || !out_docs || !out_scores || !out_counts) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_searcher_search_bool: NULL argument");
  if (nq <= 0 || top_k <= 0) NRT_FAIL(NRTGPU_ERR_INVALID, "nrtgpu_searcher_search_bool: nq and top_k must be > 0");
  NRT_CUDA_TRY(cudaSetDevice(s->ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  std::lock_guard<std::mutex> g(s->mu);
  const int64_t words = nrtgpu_packed_words(nq, top_k);
  const int n_leaves = (int)s->leaves.size();
  int rc;
  if ((rc = s->records.alloc((size_t)words * n_leaves)) || (rc = s->merged.alloc((size_t)words))) return rc;
  for (int l = 0; l < n_leaves; ++l)   // every leaf runs the whole batch (LeafCollector per segment), results stay on the device
    if ((rc = nrtgpu_search_bool_packed(s->leaves[(size_t)l], clauses, n_clauses, queries, nq, top_k, total_hits_threshold, flags, limits,
                                        stream, s->records.p + (size_t)l * words))) return rc;
  // TopDocs.merge over the leaves (LazyQueueTopScoreDocCollectorManager.java:137-144)
  if ((rc = nrtgpu_merge_topk_packed(s->ctx, n_leaves, nq, top_k, s->records.p, s->merged.p, stream))) return rc;
  s->host.resize((size_t)words);
  NRT_CUDA_TRY(cudaMemcpyAsync(s->host.data(), s->merged.p, (size_t)words * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  NRT_CUDA_TRY(cudaStreamSynchronize(st));
  const int64_t n = (int64_t)nq * top_k;
  int64_t w = 2 * n + 2ll * nq; w = (w + 1) & ~1ll;
  std::memcpy(out_docs, s->host.data(), (size_t)n * 4);
  std::memcpy(out_scores, s->host.data() + n, (size_t)n * 4);
  std::memcpy(out_counts, s->host.data() + 2 * n, (size_t)nq * 4);
  if (out_relation) for (int q = 0; q < nq; ++q) out_relation[q] = (uint8_t)(s->host[(size_t)(2 * n + nq + q)] & 1);
  if (out_total_hits) std::memcpy(out_total_hits, s->host.data() + w, (size_t)nq * 8);
  return NRTGPU_OK;
}

#include "batcher.inc"

}  // extern "C"
