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
}
