/*
 * nrtgpu_jni.c -- thin JNI shim over include/nrtgpu.h (pure marshalling; every decision is behind the C ABI).
 * NOT compiled in this repository's image (no JDK / jni.h here); build on the server host with
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude jni/nrtgpu_jni.c \
 *       -Lnrtsearch_b200 -lnrtgpu -o libnrtgpu_jni.so
 * Java side: jni/java/com/yelp/nrtsearch/server/gpu/NrtGpu.java. Buffers are direct ByteBuffers (caller allocated).
 */
#include <jni.h>
#include <stdint.h>
#include "nrtgpu.h"

#define ADDR(env, buf) ((buf) ? (*(env))->GetDirectBufferAddress((env), (buf)) : NULL)

static jint fail(JNIEnv* env, int rc) {
  if (rc == NRTGPU_OK) return 0;
  const char* cls = rc == NRTGPU_ERR_INVALID ? "java/lang/IllegalArgumentException"
                  : rc == NRTGPU_ERR_UNSUPPORTED ? "java/lang/UnsupportedOperationException"
                  : rc == NRTGPU_ERR_TIMEOUT ? "com/yelp/nrtsearch/server/search/collectors/CollectionTimeoutException"
                  : "java/lang/RuntimeException";   /* -> Status.INTERNAL in SearchHandler.handle (:136-145) */
  (*env)->ThrowNew(env, (*env)->FindClass(env, cls), nrtgpu_last_error());
  return rc;
}

JNIEXPORT jlong JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_init(JNIEnv* env, jclass c, jint device) {
  nrtgpu_ctx* ctx = NULL;
  if (fail(env, nrtgpu_init(device, &ctx))) return 0;
  return (jlong)(intptr_t)ctx;
}
JNIEXPORT void JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_shutdown(JNIEnv* env, jclass c, jlong ctx) {
  nrtgpu_shutdown((nrtgpu_ctx*)(intptr_t)ctx);
}
/* desc: a direct ByteBuffer laid out as nrtgpu_shard_desc (pointers = addresses of other direct buffers) */
JNIEXPORT jlong JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_indexBuild(JNIEnv* env, jclass c, jlong ctx, jobject desc) {
  nrtgpu_index* ix = NULL;
  if (fail(env, nrtgpu_index_build((nrtgpu_ctx*)(intptr_t)ctx, (const nrtgpu_shard_desc*)ADDR(env, desc), &ix))) return 0;
  return (jlong)(intptr_t)ix;
}
JNIEXPORT void JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_indexClose(JNIEnv* env, jclass c, jlong ix) {
  nrtgpu_index_close((nrtgpu_index*)(intptr_t)ix);
}
JNIEXPORT jint JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_searchBool(
    JNIEnv* env, jclass c, jlong ix, jobject clauses, jint nClauses, jobject queries, jint nq, jint topK,
    jint totalHitsThreshold, jint flags, jobject outDocs, jobject outScores, jobject outCounts, jobject outTotalHits,
    jobject outRelation) {
  return fail(env, nrtgpu_search_bool((nrtgpu_index*)(intptr_t)ix, (const nrtgpu_clause*)ADDR(env, clauses), nClauses,
                                      (const nrtgpu_query*)ADDR(env, queries), nq, topK, totalHitsThreshold, flags, NULL,
                                      (int32_t*)ADDR(env, outDocs), (float*)ADDR(env, outScores), (int32_t*)ADDR(env, outCounts),
                                      (int64_t*)ADDR(env, outTotalHits), (uint8_t*)ADDR(env, outRelation)));
}
JNIEXPORT jint JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_searchKnn(
    JNIEnv* env, jclass c, jlong ix, jobject queries, jint nq, jint k, jobject boosts, jobject filter, jobject outDocs,
    jobject outScores, jobject outCounts) {
  return fail(env, nrtgpu_search_knn((nrtgpu_index*)(intptr_t)ix, (const float*)ADDR(env, queries), nq, k,
                                     (const float*)ADDR(env, boosts), (const uint8_t*)ADDR(env, filter), NULL,
                                     (int32_t*)ADDR(env, outDocs), (float*)ADDR(env, outScores), (int32_t*)ADDR(env, outCounts)));
}
JNIEXPORT jint JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_blendRrf(
    JNIEnv* env, jclass c, jlong ctx, jint nRetrievers, jint nq, jint topIn, jobject docs, jobject counts, jobject boosts,
    jint rankConstant, jint topOut, jobject outDocs, jobject outScores, jobject outCounts, jobject outTotal) {
  return fail(env, nrtgpu_blend_rrf((nrtgpu_ctx*)(intptr_t)ctx, nRetrievers, nq, topIn, (const int32_t*)ADDR(env, docs),
                                    (const int32_t*)ADDR(env, counts), (const float*)ADDR(env, boosts), rankConstant, topOut,
                                    (int32_t*)ADDR(env, outDocs), (float*)ADDR(env, outScores), (int32_t*)ADDR(env, outCounts),
                                    (int32_t*)ADDR(env, outTotal)));
}
JNIEXPORT jint JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_rescoreCombine(
    JNIEnv* env, jclass c, jlong ctx, jint nq, jint nHits, jobject counts, jobject docs, jobject scores, jobject secondMatches,
    jobject secondScores, jdouble queryWeight, jdouble rescoreWeight) {
  return fail(env, nrtgpu_rescore_combine((nrtgpu_ctx*)(intptr_t)ctx, nq, nHits, (const int32_t*)ADDR(env, counts),
                                          (int32_t*)ADDR(env, docs), (float*)ADDR(env, scores),
                                          (const uint8_t*)ADDR(env, secondMatches), (const float*)ADDR(env, secondScores),
                                          queryWeight, rescoreWeight));
}

/* limits: a direct ByteBuffer laid out as nrtgpu_search_limits (or null); sort: nrtgpu_sort */
JNIEXPORT jint JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_searchBoolEx(
    JNIEnv* env, jclass c, jlong ix, jobject clauses, jint nClauses, jobject queries, jint nq, jint topK,
    jint totalHitsThreshold, jint flags, jobject limits, jobject outDocs, jobject outScores, jobject outCounts,
    jobject outTotalHits, jobject outRelation, jobject outHitTimeout, jobject outTerminatedEarly) {
  return fail(env, nrtgpu_search_bool_ex((nrtgpu_index*)(intptr_t)ix, (const nrtgpu_clause*)ADDR(env, clauses), nClauses,
                                         (const nrtgpu_query*)ADDR(env, queries), nq, topK, totalHitsThreshold, flags,
                                         (const nrtgpu_search_limits*)ADDR(env, limits), NULL, (int32_t*)ADDR(env, outDocs),
                                         (float*)ADDR(env, outScores), (int32_t*)ADDR(env, outCounts),
                                         (int64_t*)ADDR(env, outTotalHits), (uint8_t*)ADDR(env, outRelation),
                                         (uint8_t*)ADDR(env, outHitTimeout), (uint8_t*)ADDR(env, outTerminatedEarly)));
}
JNIEXPORT jint JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_searchSorted(
    JNIEnv* env, jclass c, jlong ix, jobject clauses, jint nClauses, jobject queries, jint nq, jint topK, jint flags,
    jobject sort, jobject limits, jobject outDocs, jobject outSortValues, jobject outCounts, jobject outTotalHits,
    jobject outRelation, jobject outHitTimeout, jobject outTerminatedEarly) {
  return fail(env, nrtgpu_search_sorted((nrtgpu_index*)(intptr_t)ix, (const nrtgpu_clause*)ADDR(env, clauses), nClauses,
                                        (const nrtgpu_query*)ADDR(env, queries), nq, topK, flags,
                                        (const nrtgpu_sort*)ADDR(env, sort), (const nrtgpu_search_limits*)ADDR(env, limits), NULL,
                                        (int32_t*)ADDR(env, outDocs), (int64_t*)ADDR(env, outSortValues),
                                        (int32_t*)ADDR(env, outCounts), (int64_t*)ADDR(env, outTotalHits),
                                        (uint8_t*)ADDR(env, outRelation), (uint8_t*)ADDR(env, outHitTimeout),
                                        (uint8_t*)ADDR(env, outTerminatedEarly)));
}
JNIEXPORT jint JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_scoreDocs(
    JNIEnv* env, jclass c, jlong ix, jobject clauses, jint nClauses, jobject queries, jint nq, jint nHits, jobject docs,
    jobject counts, jobject outMatches, jobject outScores) {
  return fail(env, nrtgpu_score_docs((nrtgpu_index*)(intptr_t)ix, (const nrtgpu_clause*)ADDR(env, clauses), nClauses,
                                     (const nrtgpu_query*)ADDR(env, queries), nq, nHits, (const int32_t*)ADDR(env, docs),
                                     (const int32_t*)ADDR(env, counts), NULL, (uint8_t*)ADDR(env, outMatches),
                                     (float*)ADDR(env, outScores)));
}
JNIEXPORT jint JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_fetchColumns(
    JNIEnv* env, jclass c, jlong ix, jobject colIds, jint nCols, jobject docs, jint n, jobject outValues, jobject outHas) {
  return fail(env, nrtgpu_fetch_columns((nrtgpu_index*)(intptr_t)ix, (const int32_t*)ADDR(env, colIds), nCols,
                                        (const int32_t*)ADDR(env, docs), n, NULL, (int64_t*)ADDR(env, outValues),
                                        (uint8_t*)ADDR(env, outHas)));
}
JNIEXPORT jint JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_indexSetLiveDocs(JNIEnv* env, jclass c, jlong ix, jobject live) {
  return fail(env, nrtgpu_index_set_live_docs((nrtgpu_index*)(intptr_t)ix, (const uint8_t*)ADDR(env, live)));
}
JNIEXPORT jint JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_indexUpdateStats(
    JNIEnv* env, jclass c, jlong ix, jobject termDf, jobject fieldDocCount, jobject fieldSumTtf) {
  return fail(env, nrtgpu_index_update_stats((nrtgpu_index*)(intptr_t)ix, (const int64_t*)ADDR(env, termDf),
                                             (const int64_t*)ADDR(env, fieldDocCount), (const int64_t*)ADDR(env, fieldSumTtf)));
}
/* micro-batcher: one per searcher version; submit blocks the calling gRPC handler thread until its batch is back.
 * diag: 24-byte direct buffer laid out as nrtgpu_diagnostics, or null */
JNIEXPORT jlong JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_batcherCreate(JNIEnv* env, jclass c, jlong ix, jint maxBatch, jint maxWaitUs) {
  nrtgpu_batcher* b = NULL;
  if (fail(env, nrtgpu_batcher_create((nrtgpu_index*)(intptr_t)ix, maxBatch, maxWaitUs, &b))) return 0;
  return (jlong)(intptr_t)b;
}
JNIEXPORT jint JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_batcherSubmit(
    JNIEnv* env, jclass c, jlong b, jobject clauses, jint nClauses, jint minShouldMatch, jint topK, jint totalHitsThreshold,
    jobject outDocs, jobject outScores, jobject outCount, jobject outTotalHits, jobject outRelation, jobject diag) {
  return fail(env, nrtgpu_batcher_submit((nrtgpu_batcher*)(intptr_t)b, (const nrtgpu_clause*)ADDR(env, clauses), nClauses,
                                         minShouldMatch, topK, totalHitsThreshold, (int32_t*)ADDR(env, outDocs),
                                         (float*)ADDR(env, outScores), (int32_t*)ADDR(env, outCount),
                                         (int64_t*)ADDR(env, outTotalHits), (uint8_t*)ADDR(env, outRelation),
                                         (nrtgpu_diagnostics*)ADDR(env, diag)));
}
JNIEXPORT void JNICALL Java_com_yelp_nrtsearch_server_gpu_NrtGpu_batcherClose(JNIEnv* env, jclass c, jlong b) {
  nrtgpu_batcher_close((nrtgpu_batcher*)(intptr_t)b);
}
