package com.yelp.nrtsearch.server.gpu;

import java.nio.ByteBuffer;

/** JNI binding of include/nrtgpu.h (one native method per C entry point; direct buffers only). */
public final class NrtGpu {
  static {
    System.loadLibrary("nrtgpu_jni");
  }

  private NrtGpu() {}

  public static native long init(int device);

  public static native void shutdown(long ctx);

  public static native long indexBuild(long ctx, ByteBuffer shardDesc);

  public static native void indexClose(long index);

  public static native int searchBool(
      long index, ByteBuffer clauses, int nClauses, ByteBuffer queries, int nq, int topK,
      int totalHitsThreshold, int flags, ByteBuffer outDocs, ByteBuffer outScores,
      ByteBuffer outCounts, ByteBuffer outTotalHits, ByteBuffer outRelation);

  public static native int searchKnn(
      long index, ByteBuffer queries, int nq, int k, ByteBuffer boosts, ByteBuffer filter,
      ByteBuffer outDocs, ByteBuffer outScores, ByteBuffer outCounts);

  public static native int blendRrf(
      long ctx, int nRetrievers, int nq, int topIn, ByteBuffer docs, ByteBuffer counts,
      ByteBuffer boosts, int rankConstant, int topOut, ByteBuffer outDocs, ByteBuffer outScores,
      ByteBuffer outCounts, ByteBuffer outTotal);

  public static native int rescoreCombine(
      long ctx, int nq, int nHits, ByteBuffer counts, ByteBuffer docs, ByteBuffer scores,
      ByteBuffer secondMatches, ByteBuffer secondScores, double queryWeight, double rescoreWeight);

  /** limits = nrtgpu_search_limits or null; outHitTimeout / outTerminatedEarly may be null. */
  public static native int searchBoolEx(
      long index, ByteBuffer clauses, int nClauses, ByteBuffer queries, int nq, int topK,
      int totalHitsThreshold, int flags, ByteBuffer limits, ByteBuffer outDocs,
      ByteBuffer outScores, ByteBuffer outCounts, ByteBuffer outTotalHits, ByteBuffer outRelation,
      ByteBuffer outHitTimeout, ByteBuffer outTerminatedEarly);

  public static native int searchSorted(
      long index, ByteBuffer clauses, int nClauses, ByteBuffer queries, int nq, int topK, int flags,
      ByteBuffer sort, ByteBuffer limits, ByteBuffer outDocs, ByteBuffer outSortValues,
      ByteBuffer outCounts, ByteBuffer outTotalHits, ByteBuffer outRelation,
      ByteBuffer outHitTimeout, ByteBuffer outTerminatedEarly);

  public static native int scoreDocs(
      long index, ByteBuffer clauses, int nClauses, ByteBuffer queries, int nq, int nHits,
      ByteBuffer docs, ByteBuffer counts, ByteBuffer outMatches, ByteBuffer outScores);

  public static native int fetchColumns(
      long index, ByteBuffer colIds, int nCols, ByteBuffer docs, int n, ByteBuffer outValues,
      ByteBuffer outHas);

  public static native int indexSetLiveDocs(long index, ByteBuffer liveDocs);

  public static native int indexUpdateStats(
      long index, ByteBuffer termDf, ByteBuffer fieldDocCount, ByteBuffer fieldSumTtf);

  public static native long batcherCreate(long index, int maxBatch, int maxWaitUs);

  /** Blocks until the batch this request rode in is back; diag = nrtgpu_diagnostics (24 bytes) or null. */
  public static native int batcherSubmit(
      long batcher, ByteBuffer clauses, int nClauses, int minShouldMatch, int topK,
      int totalHitsThreshold, ByteBuffer outDocs, ByteBuffer outScores, ByteBuffer outCount,
      ByteBuffer outTotalHits, ByteBuffer outRelation, ByteBuffer diag);

  public static native void batcherClose(long batcher);
}
