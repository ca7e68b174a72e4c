package com.yelp.nrtsearch.server.gpu;

import com.yelp.nrtsearch.server.search.MyIndexSearcher;
import java.io.IOException;
import java.util.concurrent.Executor;
import org.apache.lucene.index.IndexReader;
import org.apache.lucene.search.CollectorManager;
import org.apache.lucene.search.Query;

/**
 * Adaptor sketch (unbuilt here: no JDK/Lucene jars in the authoring image). A MyIndexSearcher subclass created at
 * ShardState.ShardSearcherFactory.newSearcher (ShardState.java:506-526). search() pattern-matches the rewritten Query
 * (flat BooleanQuery of TermQuery / IndexOrDocValuesQuery range / MatchAllDocsQuery, optional BoostQuery wrappers) and
 * the RelevanceCollector configuration; supported requests are queued on a micro-batcher that calls
 * NrtGpu.searchBool once per batch; everything else (and NRTGPU_ERR_UNSUPPORTED) falls through to super.search(),
 * i.e. Lucene. The GPU image is released in close(), called after the last ShardState.release (:406-425).
 */
public class GpuIndexSearcher extends MyIndexSearcher {
  private final long gpuIndex; // nrtgpu_index* of this reader version
  private final GpuQueryCompiler compiler; // term -> dense id dictionary built with the image

  protected GpuIndexSearcher(
      IndexReader reader, ExecutorAndSlicing slicing, long gpuIndex, GpuQueryCompiler compiler) {
    super(reader, slicing);
    this.gpuIndex = gpuIndex;
    this.compiler = compiler;
  }

  @Override
  public <C extends org.apache.lucene.search.Collector, T> T search(
      Query query, CollectorManager<C, T> collectorManager) throws IOException {
    GpuQueryCompiler.Compiled c = compiler.tryCompile(query, collectorManager);
    if (c == null) {
      return super.search(query, collectorManager); // not on the GPU path: Lucene
    }
    try {
      return c.toResult(GpuBatcher.forIndex(gpuIndex).submit(c).get()); // SearcherResult(TopDocs, ...)
    } catch (UnsupportedOperationException e) {
      return super.search(query, collectorManager);
    } catch (Exception e) {
      throw new IOException(e);
    }
  }

  /** Compiles Lucene queries to nrtgpu_clause/nrtgpu_query records (see INTEGRATION.md). */
  public interface GpuQueryCompiler {
    Compiled tryCompile(Query query, CollectorManager<?, ?> manager);

    interface Compiled {
      <T> T toResult(Object gpuTopDocs);
    }
  }

  /** Collects concurrent requests for <= ~200 us into one NrtGpu.searchBool call. */
  public abstract static class GpuBatcher {
    public static GpuBatcher forIndex(long gpuIndex) {
      throw new UnsupportedOperationException("sketch");
    }

    public abstract java.util.concurrent.Future<Object> submit(GpuQueryCompiler.Compiled c);
  }

  interface ExecutorAndSlicing extends Executor {}
}
