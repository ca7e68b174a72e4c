/*
 * nrtgpu.h -- C ABI of the B200-native query-execution engine that drops in behind nrtsearch's
 * SearchHandler / SearchRequestProcessor (reference = Yelp/nrtsearch @ 59c38655, Lucene 10.4.0).
 *
 * The reference has NO native seam for query execution (it is pure Java over lucene-core); the
 * entry points below are what a JNI shim would bind at the three Lucene API call sites that bound the
 * hot path:
 *
 *   nrtgpu_index_build / _close   <->  ShardSearcherFactory.newSearcher(reader, previous)
 *                                      src/main/java/com/yelp/nrtsearch/server/index/ShardState.java:506-526
 *                                      (one device image per reader version; freed after the last
 *                                      ShardState.release, :406-425)
 *   nrtgpu_search_bool            <->  searcher.search(query, collectorManager)
 *                                      src/main/java/com/yelp/nrtsearch/server/handler/SearchHandler.java:1412-1413, :556
 *                                      (BooleanQuery of TermQuery / range / match-all clauses built at
 *                                      .../query/QueryNodeMapper.java:257-283; collector config from
 *                                      .../search/collectors/RelevanceCollector.java:42-69)
 *   nrtgpu_search_knn             <->  knnQuery.rewrite(searcher)   .../search/KnnUtils.java:56
 *                                      and ExactVectorQuery          .../query/vector/ExactVectorQuery.java:137-173
 *   nrtgpu_merge_topk             <->  TopDocs.merge(0, numHits, perSlice[])
 *                                      src/main/java/org/apache/lucene/search/LazyQueueTopScoreDocCollectorManager.java:137-144
 *   nrtgpu_blend_rrf              <->  BlenderOperation.blend (weighted RRF)
 *                                      .../search/multiretriever/blender/BlenderOperation.java:76-87
 *   nrtgpu_rescore_combine        <->  RescoreOperation.rescore / QueryRescore.combine
 *                                      .../rescore/QueryRescore.java:39-57
 *
 * Conventions: every function returns 0 on success, non-zero nrtgpu_status otherwise; the message is
 * available from nrtgpu_last_error() (thread-local). Handles are opaque; output buffers are caller
 * allocated (Java direct ByteBuffers); the library never frees caller memory. All entry points are
 * re-entrant (called concurrently from the reference's SERVER / SEARCH / RETRIEVER pools).
 * There is NO CPU fallback: without a CUDA device every call fails with NRTGPU_ERR_CUDA.
 */
#ifndef NRTGPU_H
#define NRTGPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  NRTGPU_OK = 0,
  NRTGPU_ERR_INVALID = 1,      /* IllegalArgumentException on the Java side */
  NRTGPU_ERR_CUDA = 2,         /* -> Status.INTERNAL (SearchHandler.java:136-145) */
  NRTGPU_ERR_UNSUPPORTED = 3,  /* query shape outside the GPU path: caller falls through to Lucene */
  NRTGPU_ERR_OOM = 4,
  NRTGPU_ERR_TIMEOUT = 5       /* CollectionTimeoutException (SearchCutoffWrapper.java:164-174, noPartialResults) */
} nrtgpu_status;

typedef struct nrtgpu_ctx nrtgpu_ctx;     /* one per (process, device) */
typedef struct nrtgpu_index nrtgpu_index; /* device image of one shard at one reader version */
typedef struct nrtgpu_batch nrtgpu_batch; /* a compiled query batch resident on the device */

const char* nrtgpu_last_error(void);
int nrtgpu_version(void);

/* device_id: CUDA ordinal (one shard group per GPU; one process per GPU). */
int nrtgpu_init(int device_id, nrtgpu_ctx** out);
void nrtgpu_shutdown(nrtgpu_ctx* ctx);

/* BooleanClause.Occur and leaf kinds */
enum { NRTGPU_SHOULD = 0, NRTGPU_MUST = 1, NRTGPU_FILTER = 2, NRTGPU_MUST_NOT = 3 };
enum { NRTGPU_TERM = 0, NRTGPU_RANGE_I64 = 1, NRTGPU_MATCH_ALL = 2 };
/* VectorSimilarityFunction (reference VectorFieldDef.java:77-88) */
enum { NRTGPU_SIM_L2 = 0, NRTGPU_SIM_DOT = 1, NRTGPU_SIM_COSINE = 2, NRTGPU_SIM_MIP = 3 };
enum { NRTGPU_VEC_FLOAT32 = 0, NRTGPU_VEC_INT8 = 1 };

/* Host-side description of one shard, i.e. what the adaptor reads out of the LeafReaders
 * (terms()/postings()/getNormValues()/getNumericDocValues()/getFloatVectorValues()). Term ids are the
 * adaptor's dense numbering of (field, term); doc ids are shard-local, results carry doc_base + local. */
typedef struct {
  int32_t n_docs;
  int32_t doc_base;
  int32_t n_terms;
  const int64_t* term_off;          /* [n_terms+1] CSR offsets into post_* */
  const int32_t* post_docs;         /* ascending per term */
  const int32_t* post_freqs;        /* >= 1 */
  const int32_t* term_field;        /* [n_terms] text-field id, NULL = all field 0 */
  const int64_t* term_df;           /* [n_terms] INDEX-WIDE docFreq (termStatistics), NULL = CSR length */
  int32_t n_fields;
  const uint8_t* const* norms;      /* [n_fields] -> [n_docs] SmallFloat norm bytes, NULL entry = omitNorms */
  const int64_t* field_doc_count;   /* [n_fields] INDEX-WIDE collectionStatistics.docCount */
  const int64_t* field_sum_ttf;     /* [n_fields] INDEX-WIDE sumTotalTermFreq */
  const float* field_k1;            /* [n_fields] or NULL => 1.2 */
  const float* field_b;             /* [n_fields] or NULL => 0.75 */
  int32_t n_columns;
  const int64_t* const* columns;    /* [n_columns] -> [n_docs] numeric doc values (sortable-long domain) */
  const uint8_t* const* column_has; /* [n_columns] -> [n_docs] 0/1, NULL entry = every doc has a value */
  const uint8_t* live_docs;         /* [n_docs] 0/1 or NULL */
  /* one float vector field (more via nrtgpu_index_add_vectors) */
  int32_t vec_dims;                 /* 0 = none */
  int32_t vec_similarity;
  int32_t vec_count;                /* number of vectors (ord -> doc via vec_docs, NULL = identity) */
  const float* vectors;             /* [vec_count * vec_dims] float32, or int8 when vec_element_type == NRTGPU_VEC_INT8 */
  const int32_t* vec_docs;
  int32_t vec_element_type;         /* NRTGPU_VEC_FLOAT32 (FloatVectorFieldDef) or NRTGPU_VEC_INT8 (ByteVectorFieldDef: scores by
                                       VectorFieldDef.java:870-881, i.e. DOT_PRODUCT = 0.5 + dot / (dims * 2^15); queries are passed
                                       as floats holding the byte values) */
  const int64_t* const* column_offsets; /* NULL, or [n_columns]: a non-NULL entry makes column c MULTI-valued (SORTED_NUMERIC doc
                                       values, reference NumberFieldDef.java multiValued): int64[n_docs+1] offsets into columns[c],
                                       which then holds the flattened values, ascending within a doc. A range clause matches a doc
                                       when ANY of its values lies in [lo, hi] (SortedNumericDocValuesRangeQuery). Sorting, terms /
                                       min / max / sum collectors and fetch on such a column answer NRTGPU_ERR_UNSUPPORTED. */
} nrtgpu_shard_desc;

int nrtgpu_index_build(nrtgpu_ctx* ctx, const nrtgpu_shard_desc* desc, nrtgpu_index** out);
int nrtgpu_index_close(nrtgpu_index* ix);
/* bytes of device memory held by the image */
int64_t nrtgpu_index_device_bytes(const nrtgpu_index* ix);

typedef struct {
  int32_t occur;  /* NRTGPU_SHOULD.. */
  int32_t kind;   /* NRTGPU_TERM.. */
  int32_t id;     /* term id or column id */
  float boost;    /* product of enclosing BoostQuery boosts (weight = boost * idf) */
  int64_t lo, hi; /* inclusive range bounds */
} nrtgpu_clause;

typedef struct {
  int32_t clause_begin, clause_end; /* flat BooleanQuery = clauses[clause_begin:clause_end] */
  int32_t min_should_match;
  int32_t has_after;                /* searchAfter (LazyQueueTopScoreDocCollector.java:112) */
  int32_t after_doc;
  float after_score;
} nrtgpu_query;

/* flags */
enum {
  NRTGPU_FLAG_NONE = 0,
  NRTGPU_FLAG_NO_PRUNING = 1     /* force exhaustive evaluation (exact totalHits) even when total_hits_threshold < MAX
                                    would allow MAXSCORE: by default, once a query has collected more than
                                    total_hits_threshold hits, lists whose score bounds cannot reach the running k-th
                                    score stop driving the sweep -- same (doc, score) lists, totalHits becomes a lower
                                    bound (relation 1), exactly the contract of the reference's TOP_SCORES mode */
};

/* One-shot search with HOST buffers (the JNI entry point): uploads the batch, runs, copies results
 * back, synchronises `stream` (a cudaStream_t, NULL = default stream).
 *   out_docs/out_scores [nq*top_k] (score desc, doc asc), out_counts [nq],
 *   out_total_hits [nq], out_relation [nq] (0 = EQUAL_TO, 1 = GREATER_THAN_OR_EQUAL_TO).
 * total_hits_threshold == INT32_MAX <=> ScoreMode.COMPLETE (exact counts, no pruning). */
int nrtgpu_search_bool(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                       const nrtgpu_query* queries, int32_t nq, int32_t top_k,
                       int32_t total_hits_threshold, int32_t flags, void* stream, int32_t* out_docs,
                       float* out_scores, int32_t* out_counts, int64_t* out_total_hits,
                       uint8_t* out_relation);

/* Deadline and terminateAfter of a search (DocCollector config: timeoutSec, terminateAfter, terminateAfterMaxRecallCount,
 * disallowPartialResults -- reference src/main/java/com/yelp/nrtsearch/server/search/collectors/CollectorCreatorContext.java:36-53,
 * wrappers SearchCutoffWrapper.java:164-202 and TerminateAfterWrapper.java:85-162).
 *  - timeout_sec > 0: the timer starts when the first work item of the batch starts on the device (elapsed_sec = time the
 *    request already spent before the call is subtracted); it is checked at every work-item boundary ((query, <= 512K-doc
 *    slice): the reference checks per segment and, optionally, every timeoutCheckEvery docs). Work items claimed after the
 *    deadline are skipped: the query returns the hits collected so far with out_hit_timeout = 1 and relation GTE, or, with
 *    disallow_partial_results, the call fails with NRTGPU_ERR_TIMEOUT ("Search collection exceeded timeout of ...s").
 *  - terminate_after > 0: a query that has collected that many hits takes no further work items; whenever more than
 *    terminate_after docs match, out_terminated_early = 1, relation GTE and totalHits = hits counted, capped at
 *    terminate_after_max_recall_count. (The reference's slices race on one AtomicInteger, so WHICH docs are collected
 *    before the cut is timing dependent there too; here the cut falls on work-item boundaries.)
 * Both apply to queries the posting-probe kernel runs (<= 4 term clauses led by a posting list); others ignore them. */
typedef struct {
  double timeout_sec;                       /* 0: none */
  double elapsed_sec;                       /* already spent by the request before this call */
  int32_t disallow_partial_results;
  int32_t terminate_after;                  /* 0: none */
  int32_t terminate_after_max_recall_count; /* 0: = terminate_after */
} nrtgpu_search_limits;

/* nrtgpu_search_bool + limits; out_hit_timeout / out_terminated_early [nq] may be NULL */
int nrtgpu_search_bool_ex(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                          const nrtgpu_query* queries, int32_t nq, int32_t top_k,
                          int32_t total_hits_threshold, int32_t flags, const nrtgpu_search_limits* limits,
                          void* stream, int32_t* out_docs, float* out_scores, int32_t* out_counts,
                          int64_t* out_total_hits, uint8_t* out_relation, uint8_t* out_hit_timeout,
                          uint8_t* out_terminated_early);
/* same search with HOST query buffers, results left on the DEVICE in one packed record (see nrtgpu_packed_words): the
 * multi-GPU request path (the caller all-gathers the record on `stream`, then nrtgpu_merge_topk_packed). Synchronises. */
int nrtgpu_search_bool_packed(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                              const nrtgpu_query* queries, int32_t nq, int32_t top_k,
                              int32_t total_hits_threshold, int32_t flags, const nrtgpu_search_limits* limits,
                              void* stream, int32_t* d_record);

/* Sort-by-field top-k (TopFieldCollector; reference src/main/java/com/yelp/nrtsearch/server/search/collectors/
 * SortFieldCollector.java:44-105, sort construction .../search/sort/SortParser.java:54-131, numeric sort fields
 * .../field/NumberFieldDef.java:266-278 = SortedNumericSortField(type, reverse) with missingValue from
 * getSortMissingValue(missingLast)). One sort key + the implicit doc-id tie-break (lower doc first), i.e. the Sort
 * [<numeric doc-value field>], [docid] or [docid reverse]; other sorts (several fields, score mixed in) return
 * NRTGPU_ERR_UNSUPPORTED. Values live in the column's sortable-long domain (the adaptor maps int/long directly and
 * float/double through NumericUtils.floatToSortableInt / doubleToSortableLong, exactly as the range query bounds).
 *   missing_value: what a doc WITHOUT a value sorts as (Integer/Long.MIN|MAX_VALUE, -+Infinity in the sortable domain:
 *                  the reference picks MAX when missingLast, irrespective of `reverse`);
 *   after_values:  per query, the sort value of the last hit of the previous page (FieldDoc.fields[0]); used with
 *                  nrtgpu_query.has_after / after_doc (LastHitInfo, SortParser.parseLastHitInfo :131-160).
 * Results: docs in sort order, out_sort_values = the FieldDoc value of every hit (missing docs carry missing_value),
 * scores are NaN (TopFieldCollector does not track scores), totalHits exact (relation EQUAL_TO). */
enum { NRTGPU_SORT_RELEVANCE = 0, NRTGPU_SORT_COLUMN = 1, NRTGPU_SORT_DOCID = 2 };
typedef struct {
  int32_t kind;          /* NRTGPU_SORT_* */
  int32_t column;        /* NRTGPU_SORT_COLUMN: doc-value column id */
  int32_t reverse;       /* SortType.reverse */
  int32_t reserved;
  int64_t missing_value;
  const int64_t* after_values; /* [nq] or NULL */
} nrtgpu_sort;

int nrtgpu_search_sorted(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                         const nrtgpu_query* queries, int32_t nq, int32_t top_k, int32_t flags,
                         const nrtgpu_sort* sort, const nrtgpu_search_limits* limits, void* stream,
                         int32_t* out_docs, int64_t* out_sort_values, int32_t* out_counts,
                         int64_t* out_total_hits, uint8_t* out_relation, uint8_t* out_hit_timeout,
                         uint8_t* out_terminated_early);

/* Aggregating "additional collectors" over ALL docs matching each query (ScoreMode.COMPLETE: RelevanceCollector.java:55-62
 * forces totalHitsThreshold = MAX when additional collectors exist; fan-out SearchCollectorManager.java:192-198):
 *   NRTGPU_AGG_TERMS  counts per distinct value of a numeric doc-value column, the `size` buckets with the largest
 *                     (order_desc) or smallest counts, totalBuckets, totalOtherCounts
 *                     ({Int,Long,Float,Double}TermsCollectorManager + TermsCollectorManager.fillBucketResultByCount);
 *   NRTGPU_AGG_MIN / _MAX / _SUM   over the column's values as doubles (Min/Max/SumCollectorManager with the value
 *                     source doc['field'].value; no matching doc => Double.MAX_VALUE / -Double.MAX_VALUE / 0.0).
 * value_type says how the column's sortable long maps back to the number: 0 int / long, 1 float, 2 double.
 * Docs without a value contribute nothing. Bucket ties at the cut are unordered in the reference (hash-map order); here
 * the smaller value wins. Float / double sums are accumulated in a different order than the reference's single
 * thread: equal within 1e-12 relative; int / long sums below 2^53, min, max and every count are exact. */
enum { NRTGPU_AGG_TERMS = 1, NRTGPU_AGG_MIN = 2, NRTGPU_AGG_MAX = 3, NRTGPU_AGG_SUM = 4 };
typedef struct {
  int32_t kind, column, value_type;
  int32_t size;        /* terms: buckets returned (<= 2048) */
  int32_t order_desc;  /* terms: 1 = largest counts first (BucketOrder DESC by count, the default) */
  int32_t reserved;
} nrtgpu_aggregation;
typedef struct {       /* caller-allocated outputs of one aggregation (unused pointers may be NULL) */
  double* values;          /* [nq]        min / max / sum */
  int64_t* bucket_keys;    /* [nq*size]   terms: column values (sortable-long domain) */
  int32_t* bucket_counts;  /* [nq*size] */
  int32_t* n_buckets;      /* [nq]        buckets filled */
  int32_t* total_buckets;  /* [nq]        BucketResult.totalBuckets */
  int64_t* other_counts;   /* [nq]        BucketResult.totalOtherCounts */
} nrtgpu_aggregation_result;
/* nrtgpu_search_bool with additional collectors: hits as usual (exact totalHits), plus the aggregations */
int nrtgpu_search_bool_aggs(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                            const nrtgpu_query* queries, int32_t nq, int32_t top_k, int32_t flags,
                            const nrtgpu_aggregation* aggs, int32_t n_aggs, const nrtgpu_aggregation_result* results,
                            void* stream, int32_t* out_docs, float* out_scores, int32_t* out_counts,
                            int64_t* out_total_hits);

/* QueryRescorer second pass (QueryRescore.java:39-57 -> Lucene QueryRescorer.rescore): query q of the batch evaluated on
 * ITS OWN hit list docs[q][0..counts[q]) (global doc ids): out_matches / out_scores [nq*n_hits]. */
int nrtgpu_score_docs(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                      const nrtgpu_query* queries, int32_t nq, int32_t n_hits, const int32_t* docs,
                      const int32_t* counts /*[nq] or NULL*/, void* stream, uint8_t* out_matches, float* out_scores);
/* The whole rescore on the device, Lucene QueryRescorer.rescore(searcher, hits, topN = windowSize) as QueryRescore.java:52-57
 * calls it: second pass over EVERY first-pass hit + QueryRescore.combine (double math -> float) + re-sort
 * (score desc, doc asc), in place; out_counts[q] = min(counts[q], window) hits are kept.
 * docs / scores [nq*n_hits] HOST buffers (in/out), n_hits <= 4096. */
int nrtgpu_rescore_query(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                         const nrtgpu_query* queries, int32_t nq, int32_t n_hits, const int32_t* counts,
                         int32_t window, double query_weight, double rescore_weight, void* stream,
                         int32_t* docs, float* scores, int32_t* out_counts /*[nq] or NULL*/);

/* Fetch phase on doc-value columns (SearchHandler.java:397-522, FillDocsTask.fetchFromDocVales / LoadedDocValues): the
 * values of n_cols columns for n hits (global doc ids): out_values / out_has [n_cols*n]. */
int nrtgpu_fetch_columns(nrtgpu_index* ix, const int32_t* col_ids, int32_t n_cols, const int32_t* docs, int32_t n,
                         void* stream, int64_t* out_values, uint8_t* out_has);

/* Split form: compile+upload once, launch many times with everything resident in HBM. */
int nrtgpu_batch_prepare(nrtgpu_index* ix, const nrtgpu_clause* clauses, int32_t n_clauses,
                         const nrtgpu_query* queries, int32_t nq, int32_t top_k,
                         int32_t total_hits_threshold, int32_t flags, nrtgpu_batch** out);
int nrtgpu_batch_run(nrtgpu_batch* b, void* stream);   /* asynchronous: kernels only */
int nrtgpu_batch_fetch(nrtgpu_batch* b, void* stream, int32_t* out_docs, float* out_scores,
                       int32_t* out_counts, int64_t* out_total_hits, uint8_t* out_relation);
int nrtgpu_batch_set_limits(nrtgpu_batch* b, const nrtgpu_search_limits* limits);   /* NULL clears; applies to later runs */
int nrtgpu_batch_fetch_ex(nrtgpu_batch* b, void* stream, int32_t* out_docs, float* out_scores,
                          int32_t* out_counts, int64_t* out_total_hits, uint8_t* out_relation,
                          uint8_t* out_hit_timeout, uint8_t* out_terminated_early);
/* device pointers of the last run's results: uint64 keys are not exposed; these are the final arrays */
int nrtgpu_batch_device_results(nrtgpu_batch* b, int32_t** d_docs, float** d_scores, int32_t** d_counts);
/* redirect the final (docs, scores, counts) of subsequent runs into caller-owned DEVICE buffers
 * ([nq*top_k], [nq*top_k], [nq]); NULLs restore the internal buffers */
int nrtgpu_batch_bind_output(nrtgpu_batch* b, int32_t* d_docs, float* d_scores, int32_t* d_counts);
/* Packed per-shard result record: the ONE buffer a multi-GPU step all-gathers (docs, scores, counts, relation /
 * terminated flags and totalHits of every query). int32 words:
 *   docs [nq*top_k] | scores [nq*top_k] (float bits) | counts [nq] | flags [nq] (bit 0: relation GTE, bit 1: terminated
 *   early) | pad to 8 bytes | totalHits [nq] int64.      nrtgpu_packed_words = record size in words.
 * nrtgpu_batch_bind_packed redirects the results of subsequent runs into a caller-owned DEVICE record (NULL restores
 * the internal buffers); nrtgpu_merge_topk_packed is TopDocs.merge over n_lists gathered records (totalHits summed,
 * relation GTE if any shard's is: LazyQueueTopScoreDocCollectorManager.java:137-144) into one record, on the device. */
int64_t nrtgpu_packed_words(int32_t nq, int32_t top_k);
int nrtgpu_batch_bind_packed(nrtgpu_batch* b, int32_t* d_record);
int nrtgpu_merge_topk_packed(nrtgpu_ctx* ctx, int32_t n_lists, int32_t nq, int32_t top_k, const int32_t* d_records,
                             int32_t* d_out_record, void* stream);

/* stats of the compiled batch: algorithmic postings (sum of df over all term clauses), kernel launches per run */
int nrtgpu_batch_stats(const nrtgpu_batch* b, int64_t* alg_postings, int32_t* launches_per_run,
                       int64_t* work_items);
/* mean duration (ms) of stage `stage` over the runs since nrtgpu_batch_reset_timing (at most the last
 * 64; CUDA events recorded on each run's stream, which must have been synchronised).
 * stage 0 = posting traversal kernel, 1 = slice merge kernel. */
int nrtgpu_batch_stage_ms(nrtgpu_batch* b, int32_t stage, float* ms);
int nrtgpu_batch_reset_timing(nrtgpu_batch* b);
int nrtgpu_batch_free(nrtgpu_batch* b);

/* Exact kNN (ExactVectorQuery / KnnFloatVectorQuery with exact semantics): HOST buffers. Deleted docs (live_docs of the
 * shard) are never hits, as through IndexSearcher's acceptDocs; `filter` is ANDed on top. Exact BY CONSTRUCTION: the
 * tensor-core candidate stage is followed by an fp64 re-score and a rank-safety certificate (every vector outside the
 * candidate list is proven, with the bf16 error bound 2^-7 |q||d|, to score below the k-th exact score); queries the
 * certificate rejects are re-run by exact evaluation of every vector (nrtgpu_knn_last_uncertified counts them). */
int nrtgpu_search_knn(nrtgpu_index* ix, const float* queries, int32_t nq, int32_t k,
                      const float* boosts /*[nq] or NULL*/, const uint8_t* filter /*[n_docs] 0/1 or NULL*/,
                      void* stream, int32_t* out_docs, float* out_scores, int32_t* out_counts);

/* same search, plus the device time (ms, CUDA events on `stream`) of its three stages:
 * stage_ms[0] candidate GEMM (tcgen05 bf16 when dims % 8 == 0), [1] per-query select, [2] exact fp64 re-score */
int nrtgpu_search_knn_timed(nrtgpu_index* ix, const float* queries, int32_t nq, int32_t k, void* stream,
                            int32_t* out_docs, float* out_scores, int32_t* out_counts, float* stage_ms);

/* number of queries of the most recent kNN call on this index that took the exact fallback */
int32_t nrtgpu_knn_last_uncertified(const nrtgpu_index* ix);

/* TopDocs.merge over `n_lists` per-shard lists resident on the DEVICE (the receive buffer of the NCCL
 * all-gather): docs/scores [n_lists][nq][top_k], counts [n_lists][nq]; outputs on the device. */
int nrtgpu_merge_topk_device(nrtgpu_ctx* ctx, int32_t n_lists, int32_t nq, int32_t top_k,
                             const int32_t* d_docs, const float* d_scores, const int32_t* d_counts,
                             int32_t* d_out_docs, float* d_out_scores, int32_t* d_out_counts, void* stream);

/* Weighted RRF blend of R retrievers' lists for nq queries (HOST buffers):
 * docs [R][nq][top_in], counts [R][nq], boosts [R]; out [nq][top_out]. */
int nrtgpu_blend_rrf(nrtgpu_ctx* ctx, int32_t n_retrievers, int32_t nq, int32_t top_in,
                     const int32_t* docs, const int32_t* counts, const float* boosts,
                     int32_t rank_constant, int32_t top_out, int32_t* out_docs, float* out_scores,
                     int32_t* out_counts, int32_t* out_total);

/* Score-order blend (WeightedScoreOrderBlenderOperation.java:50-73 with WeightedScoreDoc.java:57-77): every hit contributes
 * score * boost of its retriever; a doc found by several retrievers combines them, in retriever declaration order and in
 * float, by MAX (default), SUM or AVG (running average). scores [R][nq][top_in]; other arguments as nrtgpu_blend_rrf. */
enum { NRTGPU_BLEND_MAX = 1, NRTGPU_BLEND_SUM = 2, NRTGPU_BLEND_AVG = 3 };
int nrtgpu_blend_scores(nrtgpu_ctx* ctx, int32_t score_mode, int32_t n_retrievers, int32_t nq, int32_t top_in,
                        const int32_t* docs, const float* scores, const int32_t* counts, const float* boosts,
                        int32_t top_out, int32_t* out_docs, float* out_scores, int32_t* out_counts, int32_t* out_total);

/* QueryRescore.combine + re-sort for nq hit lists (HOST buffers, in place): [nq][n_hits]. */
int nrtgpu_rescore_combine(nrtgpu_ctx* ctx, int32_t nq, int32_t n_hits, const int32_t* counts,
                           int32_t* docs, float* scores, const uint8_t* second_matches,
                           const float* second_scores, double query_weight, double rescore_weight);

/* ---- NRT refresh without rebuilding images (ShardSearcherFactory.newSearcher(reader, previous), ShardState.java:506-526).
 * A shard is searched through ONE image per Lucene leaf (segment); the adaptor numbers (field, term) once per shard, so
 * clause ids mean the same term in every leaf. A new reader version keeps the images of the leaves it shares with the
 * previous one and builds images only for the NEW leaves (nrtgpu_index_build with the leaf's docBase); what changes for
 * the old leaves is
 *   - their liveDocs (deletes):               nrtgpu_index_set_live_docs   (NULL = no deletes)
 *   - the index-wide statistics BM25 uses:     nrtgpu_index_update_stats    (docFreq per term, docCount and
 *     sumTotalTermFreq per field; refreshes the idf inputs, the length caches and the index-time impact bounds).
 * Both wait for searches in flight on the image. nrtgpu_searcher = the leaves of one reader version: every leaf runs the
 * batch, the per-leaf pages are merged on the device (TopDocs.merge), totalHits summed, relation GTE if any leaf's is.
 * The searcher does not own the leaves. */
int nrtgpu_index_set_live_docs(nrtgpu_index* ix, const uint8_t* live_docs /*[n_docs] 0/1 or NULL*/);
int nrtgpu_index_update_stats(nrtgpu_index* ix, const int64_t* term_df /*[n_terms] or NULL = unchanged*/,
                              const int64_t* field_doc_count /*[n_fields]*/, const int64_t* field_sum_ttf /*[n_fields]*/);
typedef struct nrtgpu_searcher nrtgpu_searcher;
int nrtgpu_searcher_create(nrtgpu_ctx* ctx, nrtgpu_index* const* leaves, int32_t n_leaves, nrtgpu_searcher** out);
int nrtgpu_searcher_search_bool(nrtgpu_searcher* s, const nrtgpu_clause* clauses, int32_t n_clauses,
                                const nrtgpu_query* queries, int32_t nq, int32_t top_k,
                                int32_t total_hits_threshold, int32_t flags, const nrtgpu_search_limits* limits,
                                void* stream, int32_t* out_docs, float* out_scores, int32_t* out_counts,
                                int64_t* out_total_hits, uint8_t* out_relation);
int nrtgpu_searcher_close(nrtgpu_searcher* s);

/* Request micro-batcher: the reference's search API is ONE query per RPC (clientlib/src/main/proto/yelp/nrtsearch/
 * luceneserver.proto:164), each on its own SERVER-pool thread (GrpcServerExecutorSupplier.java:68-75). Handler threads call
 * nrtgpu_batcher_submit (blocking) with one flat BooleanQuery; a worker thread owned by the batcher groups the waiting requests
 * that share (top_k, totalHitsThreshold) into one nrtgpu_search_bool call as soon as max_batch of them wait or the oldest has
 * waited max_wait_us, and scatters the results. A request that fails compilation is re-run alone, so it cannot fail its
 * neighbours. nrtgpu_diagnostics carries what SearchResponse.Diagnostics reports per search (SearchHandler.java:261,280,321):
 * time queued, time of the batched search, and the size of the batch the request rode in. */
typedef struct nrtgpu_batcher nrtgpu_batcher;
typedef struct { double queue_ms; double search_ms; int32_t batch_size; int32_t reserved; } nrtgpu_diagnostics;
int nrtgpu_batcher_create(nrtgpu_index* ix, int32_t max_batch, int32_t max_wait_us, nrtgpu_batcher** out);
int nrtgpu_batcher_submit(nrtgpu_batcher* b, const nrtgpu_clause* clauses, int32_t n_clauses, int32_t min_should_match,
                          int32_t top_k, int32_t total_hits_threshold, int32_t* out_docs, float* out_scores,
                          int32_t* out_count, int64_t* out_total_hits, uint8_t* out_relation,
                          nrtgpu_diagnostics* diag /* or NULL */);
int nrtgpu_batcher_stats(nrtgpu_batcher* b, int64_t* n_batches, int64_t* n_requests);
int nrtgpu_batcher_close(nrtgpu_batcher* b);   /* drains the queue, joins the worker */

#ifdef __cplusplus
}
#endif
#endif
