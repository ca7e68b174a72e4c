/*
 * oracle.c -- CPU restatement of the reference's query-execution arithmetic.
 * TEST INFRASTRUCTURE ONLY (see oracle.h). Compile: -O2 -ffp-contract=off (no FMA: Java never fuses).
 *
 * Each function cites the reference call site it follows; the Lucene 10.4.0 formulas
 * themselves are restated from the published algorithm (lucene-core is not vendored).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ SmallFloat */
/* Lucene SmallFloat.longToInt4 / int4ToLong / intToByte4 / byte4ToInt. Norms are enabled for TEXT
 * fields at reference src/main/java/com/yelp/nrtsearch/server/field/TextFieldDef.java:134. */
static int bitlen64(uint64_t x) { int n = 0; while (x) { ++n; x >>= 1; } return n; }
static int long_to_int4(uint64_t i) {
  int nbits = bitlen64(i);
  if (nbits < 4) return (int)i;
  int shift = nbits - 4;
  int enc = (int)(i >> shift) & 0x07;
  return enc | ((shift + 1) << 3);
}
static int64_t int4_to_long(int i) {
  int64_t bits = i & 0x07;
  int shift = (i >> 3) - 1;
  if (shift == -1) return bits;
  return (bits | 0x08) << shift;
}
#define NUM_FREE_VALUES 24 /* 255 - longToInt4(Integer.MAX_VALUE) = 255 - 231 */
uint8_t orc_int_to_byte4(int32_t i) {
  if (i < NUM_FREE_VALUES) return (uint8_t)i;
  return (uint8_t)(NUM_FREE_VALUES + long_to_int4((uint64_t)(i - NUM_FREE_VALUES)));
}
int32_t orc_byte4_to_int(uint8_t b) {
  int i = b;
  if (i < NUM_FREE_VALUES) return i;
  return (int32_t)(NUM_FREE_VALUES + int4_to_long(i - NUM_FREE_VALUES));
}

/* ------------------------------------------------------------------ BM25 */
/* Lucene BM25Similarity (k1=1.2, b=0.75 via `new BM25Similarity()` at reference
 * src/main/java/com/yelp/nrtsearch/server/similarity/SimilarityCreator.java:33).
 * Pinned by tests/test_oracle_golden.py against the reference's hard-coded scores. */
float orc_bm25_idf(int64_t doc_freq, int64_t doc_count) {
  return (float)log(1.0 + ((double)doc_count - (double)doc_freq + 0.5) / ((double)doc_freq + 0.5));
}
float orc_bm25_avgdl(int64_t sum_total_term_freq, int64_t doc_count) {
  return (float)((double)sum_total_term_freq / (double)doc_count);
}
void orc_bm25_cache(float k1, float b, float avgdl, float cache[256]) {
  for (int i = 0; i < 256; ++i) {
    float len = (float)orc_byte4_to_int((uint8_t)i);
    float t = b * len;
    t = t / avgdl;
    t = (1.0f - b) + t;
    t = k1 * t;
    cache[i] = 1.0f / t;
  }
}
float orc_bm25_score(float weight, float freq, uint8_t norm, const float cache[256]) {
  float x = freq * cache[norm];
  x = 1.0f + x;
  x = weight / x;
  return weight - x;
}

/* ------------------------------------------------------------------ top-k heap */
/* Order of reference src/main/java/org/apache/lucene/search/LazyQueueTopScoreDocCollector.java:129-143
 * (= Lucene HitQueue): higher score first; equal score => lower doc first. The heap root is the WORST
 * kept hit. */
/* k: sort-by-field key (0 in relevance mode): larger k ranks first, then the relevance order (TopFieldCollector: sort value,
 * then doc id -- scores are all 0 there) */
typedef struct { float score; int32_t doc; uint64_t k; } hit_t;
static inline int hit_worse(hit_t a, hit_t b) { /* a ranks after b */
  if (a.k != b.k) return a.k < b.k;
  return a.score < b.score || (a.score == b.score && a.doc > b.doc);
}
typedef struct { hit_t* h; int n, cap; } heap_t;
static void heap_sift_down(heap_t* q, int i) {
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m = i;
    if (l < q->n && hit_worse(q->h[l], q->h[m])) m = l;
    if (r < q->n && hit_worse(q->h[r], q->h[m])) m = r;
    if (m == i) return;
    hit_t t = q->h[i]; q->h[i] = q->h[m]; q->h[m] = t; i = m;
  }
}
static void heap_push(heap_t* q, hit_t x) {
  int i = q->n++;
  q->h[i] = x;
  while (i > 0) {
    int p = (i - 1) / 2;
    if (!hit_worse(q->h[i], q->h[p])) break;
    hit_t t = q->h[i]; q->h[i] = q->h[p]; q->h[p] = t; i = p;
  }
}
static int hit_cmp_best_first(const void* a, const void* b) {
  const hit_t* x = (const hit_t*)a; const hit_t* y = (const hit_t*)b;
  if (x->k != y->k) return x->k > y->k ? -1 : 1;
  if (x->score > y->score) return -1;
  if (x->score < y->score) return 1;
  return (x->doc > y->doc) - (x->doc < y->doc);
}

/* collector state per query (LazyQueueTopScoreDocCollector.collect, :103-144) */
typedef struct {
  heap_t pq; int top_k; int64_t total_hits;
  int has_after; float after_score; int32_t after_doc;
  /* TerminateAfterWrapper (reference src/main/java/com/yelp/nrtsearch/server/search/TerminateAfterWrapper.java:150-162),
   * single-threaded: docs arrive in doc order, the first terminate_after are collected, later ones are only counted,
   * up to terminate_after_max_recall_count, and set terminatedEarly */
  int64_t terminate_after, max_recall, seen; int terminated_early;
  /* sort-by-field (reference SortFieldCollector.java:44-105 -> Lucene TopFieldCollector): sort_kind 1 = numeric doc-value
   * column, 2 = doc id; a doc without a value sorts as missing_value (NumberFieldDef.java:275-276) */
  int sort_kind, sort_reverse; const int64_t* sort_col; const uint8_t* sort_has; int64_t sort_missing; uint64_t after_k;
  uint8_t* match_bitmap;   /* optional: every collected doc is flagged (what the additional collectors see) */
} collector_t;
static inline uint64_t sort_k(const collector_t* c, int64_t v) {
  uint64_t u = (uint64_t)v ^ 0x8000000000000000ull;
  return c->sort_reverse ? u : ~u;
}
static inline void collect(collector_t* c, int32_t doc, float score) {
  if (c->match_bitmap) c->match_bitmap[doc] = 1;
  if (c->terminate_after > 0) {
    int64_t cur = ++c->seen;
    if (cur > c->terminate_after) {
      c->terminated_early = 1;
      if (cur > c->max_recall) return;   /* CollectionTerminatedException */
      c->total_hits++;                  /* docCount++ without collecting */
      return;
    }
  }
  c->total_hits++;
  uint64_t k = 0;
  if (c->sort_kind) {
    int64_t v = c->sort_kind == 2 ? (int64_t)doc : ((c->sort_has && !c->sort_has[doc]) ? c->sort_missing : c->sort_col[doc]);
    k = sort_k(c, v); score = 0.0f;
    if (c->has_after && (k > c->after_k || (k == c->after_k && doc <= c->after_doc))) return;   /* FieldDoc searchAfter */
  } else if (c->has_after && (score > c->after_score || (score == c->after_score && doc <= c->after_doc))) return;
  if (c->pq.n < c->top_k) { hit_t h = {score, doc, k}; heap_push(&c->pq, h); return; }
  hit_t x = {score, doc, k};
  if (hit_worse(c->pq.h[0], x)) { c->pq.h[0] = x; heap_sift_down(&c->pq, 0); }
}
static inline float collector_min_competitive(const collector_t* c) { /* -inf until the queue is full */
  return c->pq.n < c->top_k ? -INFINITY : c->pq.h[0].score;
}

/* ------------------------------------------------------------------ boolean search */
#define WIN 4096

typedef struct {
  int occur, kind;
  /* TERM */
  const int32_t* docs; const int32_t* freqs; int64_t n, cur;
  float weight; const float* cache; const uint8_t* norms;
  float max_score; /* upper bound over the whole list (pruned mode) */
  /* RANGE */
  const int64_t* col; const uint8_t* has; const int64_t* mv_off; int64_t lo, hi;
  float const_score;
} cl_t;

/* numeric range on one doc. Single-valued column: the value lies in [lo, hi] (IntFieldDef.java:124-158, inclusive
 * bounds after the exclusive ones were stepped). Multi-valued column (SORTED_NUMERIC; Lucene
 * SortedNumericDocValuesRangeQuery): ANY of the doc's values does; a doc with no value never matches. */
static int range_clause_matches(const cl_t* c, int64_t d) {
  if (c->mv_off) {
    for (int64_t p = c->mv_off[d]; p < c->mv_off[d + 1]; ++p) if (c->col[p] >= c->lo && c->col[p] <= c->hi) return 1;
    return 0;
  }
  if (c->has && !c->has[d]) return 0;
  return c->col[d] >= c->lo && c->col[d] <= c->hi;
}

static inline float clause_term_score(const cl_t* c, int32_t doc, int32_t freq) {
  uint8_t nb = c->norms ? c->norms[doc] : 1;
  return orc_bm25_score(c->weight, (float)freq, nb, c->cache);
}

/* final score combination: Lucene BooleanScorerSupplier (pure conjunction / pure disjunction sum in
 * double; mix = ReqOptSumScorer float add when minShouldMatch==0, ConjunctionScorer double add otherwise) */
static inline float combine_score(int n_req, int msm, double must_sum, double should_sum, int should_cnt) {
  if (n_req == 0) return (float)should_sum;
  float req = (float)must_sum;
  if (should_cnt == 0) return req;
  float opt = (float)should_sum;
  if (msm > 0) return (float)((double)req + (double)opt);
  return req + opt;
}

static int64_t lower_bound_i32(const int32_t* a, int64_t lo, int64_t hi, int32_t x) {
  while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a[m] < x) lo = m + 1; else hi = m; }
  return lo;
}

typedef struct {
  double must_sum[WIN], should_sum[WIN];
  uint16_t req_cnt[WIN], should_cnt[WIN];
  uint8_t excluded[WIN];
} window_t;

static int build_clauses(const orc_index* ix, const orc_clause* cls, const orc_query* q, cl_t* out,
                         float (*field_cache)[256], uint8_t* cache_ready) {
  int n = 0;
  for (int ci = q->clause_begin; ci < q->clause_end; ++ci, ++n) {
    const orc_clause* c = &cls[ci];
    cl_t* o = &out[n];
    memset(o, 0, sizeof(*o));
    o->occur = c->occur; o->kind = c->kind;
    if (c->kind == ORC_TERM) {
      if (c->id < 0 || c->id >= ix->n_terms) return -1;
      int f = ix->term_field ? ix->term_field[c->id] : 0;
      int64_t b0 = ix->term_off[c->id], b1 = ix->term_off[c->id + 1];
      o->docs = ix->post_docs + b0; o->freqs = ix->post_freqs + b0; o->n = b1 - b0; o->cur = 0;
      int64_t df = ix->term_df ? ix->term_df[c->id] : (b1 - b0);
      float k1 = ix->field_k1 ? ix->field_k1[f] : 1.2f, b = ix->field_b ? ix->field_b[f] : 0.75f;
      if (!cache_ready[f]) {
        orc_bm25_cache(k1, b, orc_bm25_avgdl(ix->field_sum_ttf[f], ix->field_doc_count[f]), field_cache[f]);
        cache_ready[f] = 1;
      }
      o->cache = field_cache[f];
      o->norms = ix->norms ? ix->norms[f] : NULL;
      /* BM25Scorer: weight = boost * idf (float*float) */
      o->weight = c->boost * orc_bm25_idf(df > 0 ? df : 1, ix->field_doc_count[f]);
      /* list-wide upper bound: the score is monotone in x = freq*cache[norm] (float ops are monotone),
       * so score(max x) bounds every posting; without index-time impacts fall back to freq->inf = weight */
      if (ix->term_max_x) { float x = 1.0f + ix->term_max_x[c->id]; x = o->weight / x; o->max_score = o->weight - x; }
      else o->max_score = o->weight;
    } else if (c->kind == ORC_RANGE_I64) {
      if (c->id < 0 || c->id >= ix->n_columns) return -1;
      o->col = ix->columns[c->id]; o->has = ix->column_has ? ix->column_has[c->id] : NULL;
      o->mv_off = ix->column_offsets ? ix->column_offsets[c->id] : NULL;
      o->lo = c->lo; o->hi = c->hi; o->const_score = c->boost;
    } else if (c->kind == ORC_MATCH_ALL) {
      o->const_score = c->boost;
    } else return -1;
  }
  return n;
}

static inline void window_apply(window_t* w, int i, int occur, float s) {
  switch (occur) {
    case ORC_MUST: w->must_sum[i] += (double)s; w->req_cnt[i]++; break;
    case ORC_FILTER: w->req_cnt[i]++; break;
    case ORC_SHOULD: w->should_sum[i] += (double)s; w->should_cnt[i]++; break;
    default: w->excluded[i] = 1; break;
  }
}

/* Exhaustive evaluation of one flat BooleanQuery (the shape built at reference
 * src/main/java/com/yelp/nrtsearch/server/query/QueryNodeMapper.java:257-283), window by window in
 * doc order like Lucene's BooleanScorer; every matching live doc reaches the collector. */
static void search_one_exhaustive(const orc_index* ix, cl_t* cl, int ncl, int msm, collector_t* col, window_t* w) {
  int n_req = 0, n_should = 0, dense = 0;
  for (int i = 0; i < ncl; ++i) {
    if (cl[i].occur == ORC_MUST || cl[i].occur == ORC_FILTER) n_req++;
    if (cl[i].occur == ORC_SHOULD) n_should++;
    if (cl[i].kind != ORC_TERM) dense = 1;
  }
  if (msm > n_should) return;
  int need_should = msm > 0 ? msm : (n_req == 0 ? 1 : 0);
  if (n_req == 0 && n_should == 0) return;
  for (int64_t base = 0; base < ix->n_docs; base += WIN) {
    int wn = (int)((ix->n_docs - base) < WIN ? (ix->n_docs - base) : WIN);
    int32_t wend = (int32_t)(base + wn);
    if (!dense) { /* skip windows without postings */
      int any = 0;
      for (int i = 0; i < ncl; ++i) if (cl[i].cur < cl[i].n && cl[i].docs[cl[i].cur] < wend) { any = 1; break; }
      if (!any) continue;
    }
    memset(w->must_sum, 0, sizeof(double) * wn); memset(w->should_sum, 0, sizeof(double) * wn);
    memset(w->req_cnt, 0, sizeof(uint16_t) * wn); memset(w->should_cnt, 0, sizeof(uint16_t) * wn);
    memset(w->excluded, 0, wn);
    for (int i = 0; i < ncl; ++i) {
      cl_t* c = &cl[i];
      if (c->kind == ORC_TERM) {
        int64_t p = c->cur;
        while (p < c->n && c->docs[p] < wend) {
          int32_t d = c->docs[p];
          float s = (c->occur == ORC_MUST || c->occur == ORC_SHOULD) ? clause_term_score(c, d, c->freqs[p]) : 0.0f;
          window_apply(w, (int)(d - base), c->occur, s);
          ++p;
        }
        c->cur = p;
      } else if (c->kind == ORC_RANGE_I64) {
        for (int j = 0; j < wn; ++j) {
          int64_t d = base + j;
          if (range_clause_matches(c, d)) window_apply(w, j, c->occur, c->const_score);
        }
      } else {
        for (int j = 0; j < wn; ++j) window_apply(w, j, c->occur, c->const_score);
      }
    }
    for (int j = 0; j < wn; ++j) {
      if (w->excluded[j] || w->req_cnt[j] != n_req || w->should_cnt[j] < need_should) continue;
      int32_t d = (int32_t)(base + j);
      if (ix->live_docs && !ix->live_docs[d]) continue;
      float s = combine_score(n_req, msm, w->must_sum[j], w->should_sum[j], w->should_cnt[j]);
      collect(col, d, s);
    }
  }
}

/* ---- dynamic pruning (TOP_SCORES): MAXSCORE for pure disjunctions of terms, lead-list driven
 * conjunction otherwise. Rank-safe: returns exactly the exhaustive top-k. Stands in for Lucene's
 * MaxScoreBulkScorer / BlockMaxConjunctionBulkScorer as the CPU baseline. */
#define BLK 128
typedef struct { const cl_t* c; float ub; } ms_term_t;

static void search_one_pruned_disjunction(const orc_index* ix, cl_t* cl, int ncl, collector_t* col,
                                          int64_t threshold, int* pruned_out) {
  /* sort clause indices by max score ascending (MAXSCORE order) */
  int order[64]; int n = ncl;
  for (int i = 0; i < n; ++i) order[i] = i;
  for (int i = 1; i < n; ++i) { int x = order[i], j = i - 1; while (j >= 0 && cl[order[j]].max_score > cl[x].max_score) { order[j + 1] = order[j]; --j; } order[j + 1] = x; }
  int pruned = 0;
  /* DAAT over essential lists; the essential split is recomputed only when theta changes */
  float last_theta = -INFINITY;
  int first_ess = 0;
  for (;;) {
    float theta = (col->total_hits > threshold) ? collector_min_competitive(col) : -INFINITY;
    if (theta != last_theta) {
      /* essential split: largest prefix (in ascending max-score order) whose double sum of bounds cannot
       * exceed theta on its own */
      last_theta = theta;
      first_ess = 0;
      double pre = 0.0;
      while (first_ess < n) {
        double s2 = pre + (double)cl[order[first_ess]].max_score;
        if ((float)s2 > theta) break; /* a doc only in the prefix scores <= theta: not competitive (ties lose: later doc) */
        pre = s2; ++first_ess;
      }
      if (first_ess > 0) pruned = 1;
    }
    if (first_ess == n) break; /* nothing can beat theta any more */
    /* next candidate = min doc among essential cursors */
    int32_t d = INT32_MAX;
    for (int i = first_ess; i < n; ++i) { cl_t* c = &cl[order[i]]; if (c->cur < c->n && c->docs[c->cur] < d) d = c->docs[c->cur]; }
    if (d == INT32_MAX) break;
    /* advance non-essential cursors lazily to >= d (galloping) */
    for (int i = 0; i < first_ess; ++i) {
      cl_t* c = &cl[order[i]];
      if (c->cur < c->n && c->docs[c->cur] < d) {
        int64_t step = 1, lo = c->cur, hi = c->cur + 1;
        while (hi < c->n && c->docs[hi] < d) { lo = hi; step <<= 1; hi += step; }
        if (hi > c->n) hi = c->n;
        c->cur = lower_bound_i32(c->docs, lo, hi, d);
      }
    }
    /* score in CLAUSE order (double sum) */
    double sum = 0.0; int matched = 0;
    for (int i = 0; i < n; ++i) {
      cl_t* c = &cl[i];
      if (c->cur < c->n && c->docs[c->cur] == d) { sum += (double)clause_term_score(c, d, c->freqs[c->cur]); matched++; }
    }
    for (int i = first_ess; i < n; ++i) { cl_t* c = &cl[order[i]]; if (c->cur < c->n && c->docs[c->cur] == d) c->cur++; }
    if (matched && (!ix->live_docs || ix->live_docs[d])) collect(col, d, (float)sum);
  }
  *pruned_out = pruned;
}

static void run_query(const orc_index* ix, const orc_clause* cls, const orc_query* q, int top_k,
                      int64_t threshold, int mode, window_t* w, float (*field_cache)[256], uint8_t* cache_ready,
                      int32_t* out_docs, float* out_scores, int32_t* out_count, int64_t* out_total, uint8_t* out_rel,
                      int64_t terminate_after, int64_t max_recall, uint8_t* out_terminated,
                      const orc_sort* sort, int64_t after_value, int64_t* out_values, uint8_t* match_bitmap) {
  cl_t cl[64];
  int ncl = q->clause_end - q->clause_begin;
  *out_count = 0; *out_total = 0; *out_rel = 0;
  if (ncl > 64) { *out_count = -1; return; }
  ncl = build_clauses(ix, cls, q, cl, field_cache, cache_ready);
  if (ncl < 0) { *out_count = -1; return; }
  collector_t col; memset(&col, 0, sizeof(col));
  col.pq.h = (hit_t*)malloc(sizeof(hit_t) * (size_t)(top_k > 0 ? top_k : 1)); col.pq.cap = top_k; col.top_k = top_k;
  col.has_after = q->has_after; col.after_score = q->after_score; col.after_doc = q->after_doc - ix->doc_base;
  col.terminate_after = terminate_after; col.max_recall = max_recall > terminate_after ? max_recall : terminate_after;
  col.match_bitmap = match_bitmap;
  if (match_bitmap) mode = 0;
  if (terminate_after > 0) mode = 0;   /* the wrapper sees every matching doc: exhaustive evaluation */
  if (sort && sort->kind) {
    mode = 0;
    col.sort_kind = sort->kind; col.sort_reverse = sort->reverse; col.sort_missing = sort->missing_value;
    if (sort->kind == 1) {
      if (sort->column < 0 || sort->column >= ix->n_columns) { *out_count = -1; return; }
      col.sort_col = ix->columns[sort->column]; col.sort_has = ix->column_has ? ix->column_has[sort->column] : NULL;
    }
    col.after_k = sort_k(&col, sort->kind == 2 ? (int64_t)(q->after_doc - ix->doc_base) : after_value);
  }
  int pruned = 0;
  int pure_term_disj = 1;
  for (int i = 0; i < ncl; ++i) if (cl[i].kind != ORC_TERM || cl[i].occur != ORC_SHOULD) pure_term_disj = 0;
  if (mode == 1 && pure_term_disj && q->min_should_match <= 1 && ncl > 0 && !q->has_after) {
    search_one_pruned_disjunction(ix, cl, ncl, &col, threshold, &pruned);
  } else {
    search_one_exhaustive(ix, cl, ncl, q->min_should_match, &col, w);
  }
  int n = col.pq.n;
  qsort(col.pq.h, (size_t)n, sizeof(hit_t), hit_cmp_best_first);
  for (int i = 0; i < n; ++i) {
    out_docs[i] = col.pq.h[i].doc + ix->doc_base; out_scores[i] = col.pq.h[i].score;
    if (out_values) {
      uint64_t u = col.sort_reverse ? col.pq.h[i].k : ~col.pq.h[i].k;
      out_values[i] = col.sort_kind == 2 ? (int64_t)out_docs[i] : (int64_t)(u ^ 0x8000000000000000ull);
    }
  }
  *out_count = n; *out_total = col.total_hits; *out_rel = (pruned || col.terminated_early) ? 1 : 0;
  if (out_terminated) *out_terminated = (uint8_t)col.terminated_early;
  free(col.pq.h);
}

int orc_search(const orc_index* ix, const orc_clause* clauses, const orc_query* queries, int32_t nq,
               int32_t top_k, int32_t total_hits_threshold, int32_t mode, int32_t n_threads,
               int32_t* out_docs, float* out_scores, int32_t* out_counts, int64_t* out_total,
               uint8_t* out_rel) {
  return orc_search_limits(ix, clauses, queries, nq, top_k, total_hits_threshold, mode, n_threads, 0, 0, out_docs, out_scores,
                           out_counts, out_total, out_rel, NULL);
}

static int search_all(const orc_index* ix, const orc_clause* clauses, const orc_query* queries, int32_t nq,
                      int32_t top_k, int32_t total_hits_threshold, int32_t mode, int32_t n_threads,
                      int32_t terminate_after, int32_t terminate_after_max_recall, const orc_sort* sort, const int64_t* after_values,
                      int32_t* out_docs, float* out_scores, int64_t* out_values, int32_t* out_counts, int64_t* out_total,
                      uint8_t* out_rel, uint8_t* out_terminated);

int orc_search_sorted(const orc_index* ix, const orc_clause* clauses, const orc_query* queries, int32_t nq, int32_t top_k,
                      int32_t n_threads, const orc_sort* sort, const int64_t* after_values, int32_t* out_docs,
                      int64_t* out_values, int32_t* out_counts, int64_t* out_total) {
  float* scores = (float*)malloc(sizeof(float) * (size_t)nq * (size_t)top_k);
  uint8_t* rel = (uint8_t*)malloc((size_t)nq);
  int rc = search_all(ix, clauses, queries, nq, top_k, INT32_MAX, 0, n_threads, 0, 0, sort, after_values, out_docs, scores, out_values,
                      out_counts, out_total, rel, NULL);
  free(scores); free(rel);
  return rc;
}

int orc_search_limits(const orc_index* ix, const orc_clause* clauses, const orc_query* queries, int32_t nq,
                      int32_t top_k, int32_t total_hits_threshold, int32_t mode, int32_t n_threads,
                      int32_t terminate_after, int32_t terminate_after_max_recall,
                      int32_t* out_docs, float* out_scores, int32_t* out_counts, int64_t* out_total,
                      uint8_t* out_rel, uint8_t* out_terminated) {
  return search_all(ix, clauses, queries, nq, top_k, total_hits_threshold, mode, n_threads, terminate_after, terminate_after_max_recall,
                    NULL, NULL, out_docs, out_scores, NULL, out_counts, out_total, out_rel, out_terminated);
}

static int search_all(const orc_index* ix, const orc_clause* clauses, const orc_query* queries, int32_t nq,
                      int32_t top_k, int32_t total_hits_threshold, int32_t mode, int32_t n_threads,
                      int32_t terminate_after, int32_t terminate_after_max_recall, const orc_sort* sort, const int64_t* after_values,
                      int32_t* out_docs, float* out_scores, int64_t* out_values, int32_t* out_counts, int64_t* out_total,
                      uint8_t* out_rel, uint8_t* out_terminated) {
  if (top_k <= 0) return -1;
  /* manager rule: threshold = max(threshold, numHits)
   * (reference LazyQueueTopScoreDocCollectorManager.java:102) */
  int64_t thr = total_hits_threshold < top_k ? top_k : total_hits_threshold;
  if (total_hits_threshold == INT32_MAX) mode = 0; /* ScoreMode.COMPLETE (:68-70) */
  int bad = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
  int nf = ix->n_fields > 0 ? ix->n_fields : 1;
#pragma omp parallel
  {
    window_t* w = (window_t*)malloc(sizeof(window_t));
    float (*fc)[256] = (float (*)[256])malloc(sizeof(float) * 256 * (size_t)nf);
    uint8_t* ready = (uint8_t*)calloc((size_t)nf, 1);
#pragma omp for schedule(dynamic, 1)
    for (int qi = 0; qi < nq; ++qi) {
      run_query(ix, clauses, &queries[qi], top_k, thr, mode, w, fc, ready, out_docs + (size_t)qi * top_k,
                out_scores + (size_t)qi * top_k, &out_counts[qi], &out_total[qi], &out_rel[qi],
                terminate_after, terminate_after_max_recall, out_terminated ? &out_terminated[qi] : NULL,
                sort, after_values ? after_values[qi] : 0, out_values ? out_values + (size_t)qi * top_k : NULL, NULL);
      if (out_counts[qi] < 0) {
#pragma omp atomic write
        bad = 1;
      }
    }
    free(w); free(fc); free(ready);
  }
  return bad ? -1 : 0;
}

/* every doc matching ONE query, as a 0/1 byte per doc (the doc stream the reference's additional collectors -- terms /
 * min / max / sum, SearchCollectorManager.java:192-198 -- receive) */
int orc_match_bitmap(const orc_index* ix, const orc_clause* clauses, const orc_query* query, uint8_t* out_bitmap /*[n_docs], zeroed*/) {
  window_t* w = (window_t*)malloc(sizeof(window_t));
  int nf = ix->n_fields > 0 ? ix->n_fields : 1;
  float (*fc)[256] = (float (*)[256])malloc(sizeof(float) * 256 * (size_t)nf);
  uint8_t* ready = (uint8_t*)calloc((size_t)nf, 1);
  int32_t d[1]; float s[1]; int32_t cnt; int64_t tot; uint8_t rel;
  run_query(ix, clauses, query, 1, INT32_MAX, 0, w, fc, ready, d, s, &cnt, &tot, &rel, 0, 0, NULL, NULL, 0, NULL, out_bitmap);
  free(w); free(fc); free(ready);
  return cnt < 0 ? -1 : 0;
}

/* QueryRescorer second pass: query q on its own hit list (Lucene QueryRescorer.rescore advances the second query's scorer
 * to every first-pass hit; reference src/main/java/com/yelp/nrtsearch/server/rescore/QueryRescore.java:52-57) */
int orc_score_docs(const orc_index* ix, const orc_clause* clauses, const orc_query* queries, int32_t nq, int32_t n_hits,
                   const int32_t* docs, const int32_t* counts, uint8_t* out_matches, float* out_scores) {
  int nf = ix->n_fields > 0 ? ix->n_fields : 1;
  float (*fc)[256] = (float (*)[256])malloc(sizeof(float) * 256 * (size_t)nf);
  uint8_t* ready = (uint8_t*)calloc((size_t)nf, 1);
  int bad = 0;
  for (int q = 0; q < nq && !bad; ++q) {
    cl_t cl[64];
    const orc_query* qq = &queries[q];
    int ncl = qq->clause_end - qq->clause_begin;
    if (ncl > 64) { bad = 1; break; }
    ncl = build_clauses(ix, clauses, qq, cl, fc, ready);
    if (ncl < 0) { bad = 1; break; }
    int n_req = 0, n_should = 0;
    for (int i = 0; i < ncl; ++i) {
      if (cl[i].occur == ORC_MUST || cl[i].occur == ORC_FILTER) n_req++;
      if (cl[i].occur == ORC_SHOULD) n_should++;
    }
    int msm = qq->min_should_match;
    int need_should = msm > 0 ? msm : (n_req == 0 ? 1 : 0);
    int n = counts ? counts[q] : n_hits;
    for (int h = 0; h < n_hits; ++h) {
      size_t o = (size_t)q * n_hits + h;
      out_matches[o] = 0; out_scores[o] = 0.0f;
      if (h >= n) continue;
      int64_t d = (int64_t)docs[o] - ix->doc_base;
      if (d < 0 || d >= ix->n_docs) continue;
      if (ix->live_docs && !ix->live_docs[d]) continue;
      if (msm > n_should || (n_req == 0 && n_should == 0)) continue;
      double must_sum = 0.0, should_sum = 0.0; int req_cnt = 0, should_cnt = 0, excluded = 0;
      for (int i = 0; i < ncl; ++i) {
        cl_t* c = &cl[i];
        int present = 0; float s = 0.0f;
        if (c->kind == ORC_TERM) {
          int64_t p = lower_bound_i32(c->docs, 0, c->n, (int32_t)d);
          if (p < c->n && c->docs[p] == (int32_t)d) { present = 1; if (c->occur == ORC_MUST || c->occur == ORC_SHOULD) s = clause_term_score(c, (int32_t)d, c->freqs[p]); }
        } else if (c->kind == ORC_RANGE_I64) {
          present = range_clause_matches(c, d);
          s = c->const_score;
        } else { present = 1; s = c->const_score; }
        if (!present) continue;
        switch (c->occur) {
          case ORC_MUST: must_sum += (double)s; req_cnt++; break;
          case ORC_FILTER: req_cnt++; break;
          case ORC_SHOULD: should_sum += (double)s; should_cnt++; break;
          default: excluded = 1; break;
        }
      }
      if (excluded || req_cnt != n_req || should_cnt < need_should) continue;
      out_matches[o] = 1;
      out_scores[o] = combine_score(n_req, msm, must_sum, should_sum, should_cnt);
    }
  }
  free(fc); free(ready);
  return bad ? -1 : 0;
}

/* index-time impacts: per term max of x = freq * cache[norm] (what Lucene stores as competitive
 * (freq, norm) impacts in its skip data); query-independent given avgdl. */
void orc_build_term_max_x(const orc_index* ix, float* term_max_x) {
  int nf = ix->n_fields > 0 ? ix->n_fields : 1;
  float (*fc)[256] = (float (*)[256])malloc(sizeof(float) * 256 * (size_t)nf);
  for (int f = 0; f < nf; ++f) {
    float k1 = ix->field_k1 ? ix->field_k1[f] : 1.2f, b = ix->field_b ? ix->field_b[f] : 0.75f;
    orc_bm25_cache(k1, b, orc_bm25_avgdl(ix->field_sum_ttf[f], ix->field_doc_count[f]), fc[f]);
  }
#pragma omp parallel for schedule(dynamic, 256)
  for (int t = 0; t < ix->n_terms; ++t) {
    int f = ix->term_field ? ix->term_field[t] : 0;
    const uint8_t* norms = ix->norms ? ix->norms[f] : NULL;
    float m = 0.0f;
    for (int64_t p = ix->term_off[t]; p < ix->term_off[t + 1]; ++p) {
      float x = (float)ix->post_freqs[p] * fc[f][norms ? norms[ix->post_docs[p]] : 1];
      if (x > m) m = x;
    }
    term_max_x[t] = m;
  }
  free(fc);
}

/* ------------------------------------------------------------------ TopDocs.merge */
/* reference src/main/java/org/apache/lucene/search/LazyQueueTopScoreDocCollectorManager.java:137-144:
 * TopDocs.merge(0, numHits, perSlice[]) -- score desc, then doc asc (global ids are unique). */
void orc_merge_topk(int32_t n_lists, int32_t nq, int32_t top_k, const int32_t* docs, const float* scores,
                    const int32_t* counts, int32_t* out_docs, float* out_scores, int32_t* out_counts) {
  hit_t* buf = (hit_t*)malloc(sizeof(hit_t) * (size_t)n_lists * (size_t)top_k);
  for (int q = 0; q < nq; ++q) {
    int n = 0;
    for (int l = 0; l < n_lists; ++l) {
      size_t base = ((size_t)l * nq + q) * top_k;
      for (int i = 0; i < counts[(size_t)l * nq + q]; ++i) { buf[n].doc = docs[base + i]; buf[n].score = scores[base + i]; buf[n].k = 0; ++n; }
    }
    qsort(buf, (size_t)n, sizeof(hit_t), hit_cmp_best_first);
    if (n > top_k) n = top_k;
    for (int i = 0; i < n; ++i) { out_docs[(size_t)q * top_k + i] = buf[i].doc; out_scores[(size_t)q * top_k + i] = buf[i].score; }
    out_counts[q] = n;
  }
  free(buf);
}

/* ------------------------------------------------------------------ vectors */
/* Lucene VectorSimilarityFunction.compare for float vectors; formulas restated by the reference at
 * src/main/java/com/yelp/nrtsearch/server/field/VectorFieldDef.java:664-673. Accumulation is in double
 * (Lucene's Panama float kernels are lane-order dependent, so parity is 1e-5 relative, not bitwise). */
float orc_vector_score_f32(const float* a, const float* b, int32_t dims, int32_t sim) {
  double dot = 0, na = 0, nb = 0, d2 = 0;
  for (int i = 0; i < dims; ++i) {
    double x = a[i], y = b[i];
    dot += x * y; na += x * x; nb += y * y; d2 += (x - y) * (x - y);
  }
  /* byte vectors (sim | 0x100; the values are the bytes): reference VectorFieldDef.java:870-881 -- only DOT_PRODUCT differs */
  if (sim == (ORC_SIM_DOT | 0x100)) return 0.5f + (float)dot / (float)(dims * (1 << 15));
  switch (sim & 0xff) {
    case ORC_SIM_L2: return 1.0f / (1.0f + (float)d2);
    case ORC_SIM_DOT: { float s = (1.0f + (float)dot) / 2.0f; return s > 0 ? s : 0; }
    case ORC_SIM_COSINE: { float c = (float)(dot / sqrt(na * nb)); float s = (1.0f + c) / 2.0f; return s > 0 ? s : 0; }
    case ORC_SIM_MIP: { float s = (float)dot; return s < 0 ? 1.0f / (1.0f + -1.0f * s) : s + 1.0f; }
  }
  return NAN;
}

/* reference src/main/java/com/yelp/nrtsearch/server/query/vector/ExactVectorQuery.java:137-173:
 * every doc with a vector is scored, score = vectorScorer.score() * boost, collected top-k. */
int orc_knn_exact(const float* corpus, int32_t n, int32_t dims, int32_t sim, int32_t doc_base,
                  const uint8_t* filter, const float* queries, int32_t nq, const float* boosts, int32_t k,
                  int32_t n_threads, int32_t* out_docs, float* out_scores, int32_t* out_counts,
                  const uint8_t* live_docs) {
  if (k <= 0) return -1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int q = 0; q < nq; ++q) {
    collector_t col; memset(&col, 0, sizeof(col));
    col.pq.h = (hit_t*)malloc(sizeof(hit_t) * (size_t)k); col.pq.cap = k; col.top_k = k;
    float boost = boosts ? boosts[q] : 1.0f;
    const float* qv = queries + (size_t)q * dims;
    for (int32_t d = 0; d < n; ++d) {
      if (filter && !filter[d]) continue;
      if (live_docs && !live_docs[d]) continue; /* IndexSearcher acceptDocs: deleted docs are never scored */
      float s = orc_vector_score_f32(qv, corpus + (size_t)d * dims, dims, sim) * boost;
      collect(&col, d, s);
    }
    int m = col.pq.n;
    qsort(col.pq.h, (size_t)m, sizeof(hit_t), hit_cmp_best_first);
    for (int i = 0; i < m; ++i) { out_docs[(size_t)q * k + i] = col.pq.h[i].doc + doc_base; out_scores[(size_t)q * k + i] = col.pq.h[i].score; }
    out_counts[q] = m;
    free(col.pq.h);
  }
  return 0;
}

/* ------------------------------------------------------------------ hybrid stages */
/* reference .../blender/operation/WeightedRrfBlenderOperation.java:52-78 and
 * .../blender/score/WeightedRRFScoreDoc.java:60-77: first hit score = boost/(k+rank), later
 * retrievers add boost/(k+rank) in float, in declaration order. The reference's final heap orders by
 * score only (ties unordered, BlenderOperation.java:99-132); here ties break on doc asc (a valid order). */
int orc_blend_rrf(int32_t n_retrievers, int32_t top_in, const int32_t* docs, const int32_t* counts,
                  const float* boosts, int32_t rank_constant, int32_t top_out, int32_t* out_docs,
                  float* out_scores, int32_t* total) {
  int k = rank_constant > 0 ? rank_constant : 60;
  size_t cap = (size_t)n_retrievers * (size_t)top_in;
  hit_t* m = (hit_t*)malloc(sizeof(hit_t) * (cap ? cap : 1));
  int n = 0;
  for (int r = 0; r < n_retrievers; ++r) {
    for (int i = 0; i < counts[r]; ++i) {
      int32_t d = docs[(size_t)r * top_in + i];
      float add = boosts[r] / (float)(k + (i + 1));
      int j = 0;
      for (; j < n; ++j) if (m[j].doc == d) break;
      if (j == n) { m[n].doc = d; m[n].score = add; m[n].k = 0; ++n; } else m[j].score += add;
    }
  }
  *total = n;
  qsort(m, (size_t)n, sizeof(hit_t), hit_cmp_best_first);
  int o = n < top_out ? n : top_out;
  for (int i = 0; i < o; ++i) { out_docs[i] = m[i].doc; out_scores[i] = m[i].score; }
  free(m);
  return o;
}

/* reference .../blender/operation/WeightedScoreOrderBlenderOperation.java:50-73 + .../blender/score/WeightedScoreDoc.java:57-77:
 * first hit score = score * boost; later retrievers combine score * boost by MAX (1), SUM (2) or running AVG (3), float ops */
int orc_blend_scores(int32_t mode, int32_t n_retrievers, int32_t top_in, const int32_t* docs, const float* scores,
                     const int32_t* counts, const float* boosts, int32_t top_out, int32_t* out_docs, float* out_scores, int32_t* total) {
  size_t cap = (size_t)n_retrievers * (size_t)top_in;
  hit_t* m = (hit_t*)malloc(sizeof(hit_t) * (cap ? cap : 1));
  int* have = (int*)malloc(sizeof(int) * (cap ? cap : 1));
  int n = 0;
  for (int r = 0; r < n_retrievers; ++r) {
    for (int i = 0; i < counts[r]; ++i) {
      int32_t d = docs[(size_t)r * top_in + i];
      float w = scores[(size_t)r * top_in + i] * boosts[r];
      int j = 0;
      for (; j < n; ++j) if (m[j].doc == d) break;
      if (j == n) { m[n].doc = d; m[n].score = w; m[n].k = 0; have[n] = 1; ++n; }
      else {
        if (mode == 1) m[j].score = m[j].score > w ? m[j].score : w;
        else if (mode == 2) m[j].score = m[j].score + w;
        else m[j].score = (m[j].score * (float)have[j] + w) / (float)(have[j] + 1);
        have[j]++;
      }
    }
  }
  *total = n;
  qsort(m, (size_t)n, sizeof(hit_t), hit_cmp_best_first);
  int o = n < top_out ? n : top_out;
  for (int i = 0; i < o; ++i) { out_docs[i] = m[i].doc; out_scores[i] = m[i].score; }
  free(m); free(have);
  return o;
}

/* Lucene QueryRescorer.rescore(searcher, hits, topN=windowSize) as driven by reference
 * src/main/java/com/yelp/nrtsearch/server/rescore/QueryRescore.java:39-57: every first-pass hit is
 * combined (double math -> float), re-sorted (score desc, doc asc); caller keeps the first `window`. */
void orc_rescore_combine(int32_t n_hits, int32_t window, int32_t* docs, float* scores,
                         const uint8_t* second_matches, const float* second_scores, double query_weight,
                         double rescore_weight) {
  (void)window;
  hit_t* h = (hit_t*)malloc(sizeof(hit_t) * (size_t)(n_hits ? n_hits : 1));
  for (int i = 0; i < n_hits; ++i) {
    float s = second_matches[i] ? (float)(query_weight * (double)scores[i] + rescore_weight * (double)second_scores[i])
                                : (float)(query_weight * (double)scores[i]);
    h[i].doc = docs[i]; h[i].score = s; h[i].k = 0;
  }
  qsort(h, (size_t)n_hits, sizeof(hit_t), hit_cmp_best_first);
  for (int i = 0; i < n_hits; ++i) { docs[i] = h[i].doc; scores[i] = h[i].score; }
  free(h);
}
