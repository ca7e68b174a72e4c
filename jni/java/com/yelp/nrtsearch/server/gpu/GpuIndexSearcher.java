package com.yelp.nrtsearch.server.gpu;

import com.yelp.nrtsearch.server.search.MyIndexSearcher;
import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import org.apache.lucene.index.IndexReader;
import org.apache.lucene.search.CollectorManager;
import org.apache.lucene.search.Query;
import org.apache.lucene.search.ScoreDoc;
import org.apache.lucene.search.TopDocs;
import org.apache.lucene.search.TotalHits;

/**
 * Reference-side adaptor (not built in the authoring image: no JDK / Lucene jars there). A MyIndexSearcher subclass
 * created at ShardState.ShardSearcherFactory.newSearcher (ShardState.java:506-526). search() pattern-matches the
 * rewritten Query (flat BooleanQuery of TermQuery / IndexOrDocValuesQuery range / MatchAllDocsQuery, optional
 * BoostQuery wrappers) and the RelevanceCollector configuration. A supported request is handed to the NATIVE
 * micro-batcher (nrtgpu_batcher_submit): the gRPC handler thread blocks while a worker thread inside libnrtgpu groups
 * the waiting requests into one batched search. Everything else, and NRTGPU_ERR_UNSUPPORTED
 * (UnsupportedOperationException), falls through to super.search(), i.e. Lucene. The image and its batcher are released
 * in close(), called after the last ShardState.release (:406-425).
 */
public class GpuIndexSearcher extends MyIndexSearcher {
  private final long gpuIndex; // nrtgpu_index* of this reader version
  private final long batcher; // nrtgpu_batcher* bound to gpuIndex
  private final GpuQueryCompiler compiler; // term -> dense id dictionary built with the image

  protected GpuIndexSearcher(
      IndexReader reader,
      java.util.concurrent.Executor executor,
      long gpuIndex,
      GpuQueryCompiler compiler,
      int maxBatch,
      int maxWaitUs) {
    super(reader, executor);
    this.gpuIndex = gpuIndex;
    this.compiler = compiler;
    this.batcher = NrtGpu.batcherCreate(gpuIndex, maxBatch, maxWaitUs);
  }

  @Override
  public <C extends org.apache.lucene.search.Collector, T> T search(
      Query query, CollectorManager<C, T> collectorManager) throws IOException {
    GpuQueryCompiler.Compiled c = compiler.tryCompile(query, collectorManager);
    if (c == null) {
      return super.search(query, collectorManager); // not on the GPU path: Lucene
    }
    int k = c.topK();
    ByteBuffer docs = direct(4 * k), scores = direct(4 * k), count = direct(4), total = direct(8);
    ByteBuffer relation = direct(1), diag = direct(24);
    try {
      NrtGpu.batcherSubmit(
          batcher, c.clauses(), c.numClauses(), c.minShouldMatch(), k, c.totalHitsThreshold(), docs,
          scores, count, total, relation, diag);
    } catch (UnsupportedOperationException e) {
      return super.search(query, collectorManager);
    }
    int n = count.getInt(0);
    ScoreDoc[] hits = new ScoreDoc[n];
    for (int i = 0; i < n; ++i) {
      hits[i] = new ScoreDoc(docs.getInt(4 * i), scores.getFloat(4 * i));
    }
    TotalHits.Relation rel =
        relation.get(0) == 0
            ? TotalHits.Relation.EQUAL_TO
            : TotalHits.Relation.GREATER_THAN_OR_EQUAL_TO;
    // Diagnostics of the request (SearchHandler.java:261,280,321): queue_ms, search_ms, batch_size
    return c.toResult(
        new TopDocs(new TotalHits(total.getLong(0), rel), hits),
        diag.getDouble(0),
        diag.getDouble(8),
        diag.getInt(16));
  }

  public void close() {
    NrtGpu.batcherClose(batcher);
    NrtGpu.indexClose(gpuIndex);
  }

  private static ByteBuffer direct(int bytes) {
    return ByteBuffer.allocateDirect(bytes).order(ByteOrder.nativeOrder());
  }

  /** Compiles Lucene queries to nrtgpu_clause records (see INTEGRATION.md for the rules). */
  public interface GpuQueryCompiler {
    Compiled tryCompile(Query query, CollectorManager<?, ?> manager);

    interface Compiled {
      ByteBuffer clauses(); // direct, nrtgpu_clause[numClauses]

      int numClauses();

      int minShouldMatch();

      int topK();

      int totalHitsThreshold();

      <T> T toResult(TopDocs topDocs, double queueMs, double searchMs, int batchSize);
    }
  }
}
