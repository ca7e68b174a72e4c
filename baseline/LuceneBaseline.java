/*
 * LuceneBaseline.java -- the reference's OWN search path (Lucene IndexSearcher with concurrent segment search, as
 * nrtsearch runs it: ShardState.java:506-526 MyIndexSearcher + SearchHandler.java:166-186) timed on the workload of
 * bench.py (BASELINE.json configs[1]: 10M synthetic Zipf docs, 1024 three-term disjunctive BM25 queries, top-100,
 * totalHitsThreshold 1000).
 *
 * UNVERIFIED HARNESS: the authoring image has no JDK and no Lucene jars, so this file has never been compiled here and no
 * number in this repository comes from it. bench.py's cpu_baseline / --impl reference legs time oracle/ (the C restatement
 * of the same algorithms, pinned against the reference's golden vectors) instead. A maintainer with the reference's
 * classpath can run it on the GPU box's host to replace that port figure:
 *
 *   javac -cp "$NRTSEARCH_HOME/build/install/nrtsearch/lib/*" baseline/LuceneBaseline.java -d /tmp/lb
 *   java  -cp "/tmp/lb:$NRTSEARCH_HOME/build/install/nrtsearch/lib/*" LuceneBaseline \
 *         --docs 10000000 --vocab 1000000 --queries 1024 --topk 100 --threshold 1000 --threads 128 --index /tmp/lb-index
 *
 * The corpus and the queries are the ones nrtsearch_b200/csrc/synth.cpp generates (same counter-based generator restated
 * below: mix64 / u01, Zipf(1.0) token ranks through the same guide table, doc length 8 + Poisson(56), query term ranks
 * log-uniform in [10, 10^4)), so a term's postings here equal `post_docs` there and the doc ids of the two systems can be
 * compared directly when the index is built with ONE thread and no merges reordering docs (LogByteSizeMergePolicy keeps
 * doc order; --check prints the first hits for that comparison).
 */
import java.nio.file.Paths;
import java.util.ArrayList;
import java.util.Arrays;
import java.util.List;
import java.util.concurrent.ExecutorService;
import java.util.concurrent.Executors;
import java.util.concurrent.Future;
import org.apache.lucene.analysis.core.WhitespaceAnalyzer;
import org.apache.lucene.document.Document;
import org.apache.lucene.document.Field;
import org.apache.lucene.document.TextField;
import org.apache.lucene.index.DirectoryReader;
import org.apache.lucene.index.IndexWriter;
import org.apache.lucene.index.IndexWriterConfig;
import org.apache.lucene.index.LogByteSizeMergePolicy;
import org.apache.lucene.index.Term;
import org.apache.lucene.search.BooleanClause;
import org.apache.lucene.search.BooleanQuery;
import org.apache.lucene.search.IndexSearcher;
import org.apache.lucene.search.Query;
import org.apache.lucene.search.TermQuery;
import org.apache.lucene.search.TopDocs;
import org.apache.lucene.search.TopScoreDocCollectorManager;
import org.apache.lucene.store.FSDirectory;

public final class LuceneBaseline {
  static final long SEED_CORPUS = 0x5EED0001L; // nrtsearch_b200/index.py SEED_CORPUS
  static final long SEED_QUERIES = 0x5EED0002L; // nrtsearch_b200/index.py SEED_QUERIES

  static long mix64(long z) {
    z += 0x9E3779B97F4A7C15L;
    z = (z ^ (z >>> 30)) * 0xBF58476D1CE4E5B9L;
    z = (z ^ (z >>> 27)) * 0x94D049BB133111EBL;
    return z ^ (z >>> 31);
  }

  static double u01(long seed, long counter) {
    return (double) (mix64(seed ^ mix64(counter)) >>> 11) * (1.0 / 9007199254740992.0);
  }

  /** Zipf(s) sampler with the generator's 2^20-entry guide table. */
  static final class Zipf {
    static final int GBITS = 20;
    final double[] cdf;
    final int[] guide = new int[(1 << GBITS) + 2];

    Zipf(int vocab, double s) {
      cdf = new double[vocab];
      double h = 0;
      for (int r = 0; r < vocab; ++r) {
        h += 1.0 / Math.pow(r + 1, s);
        cdf[r] = h;
      }
      for (int r = 0; r < vocab; ++r) cdf[r] /= h;
      cdf[vocab - 1] = 1.0;
      int r = 0, g = 1 << GBITS;
      for (int b = 0; b <= g; ++b) {
        double x = (double) b / (double) g;
        while (r < vocab - 1 && cdf[r] < x) ++r;
        guide[b] = r;
      }
    }

    int sample(double u) {
      int b = (int) (u * (double) (1 << GBITS));
      int lo = guide[b], hi = guide[b + 1];
      while (lo < hi) {
        int mid = (lo + hi) >>> 1;
        if (cdf[mid] > u) hi = mid;
        else lo = mid + 1;
      }
      return lo;
    }
  }

  static final class Poisson {
    final double[] cdf;

    Poisson(double lam) {
      List<Double> c = new ArrayList<>();
      double p = Math.exp(-lam), acc = p;
      c.add(acc);
      for (int k = 1; k < 1000; ++k) {
        p *= lam / k;
        acc += p;
        c.add(acc);
        if (1.0 - acc < 1e-17 && k > lam) break;
      }
      cdf = c.stream().mapToDouble(Double::doubleValue).toArray();
    }

    int sample(double u) { // first k with cdf[k] > u
      int lo = 0, hi = cdf.length;
      while (lo < hi) {
        int mid = (lo + hi) >>> 1;
        if (cdf[mid] <= u) lo = mid + 1;
        else hi = mid;
      }
      return lo;
    }
  }

  public static void main(String[] a) throws Exception {
    long docs = 10_000_000;
    int vocab = 1_000_000, nq = 1024, topk = 100, threshold = 1000, threads = Runtime.getRuntime().availableProcessors();
    int repeat = 3;
    String index = "/tmp/lb-index";
    boolean check = false;
    for (int i = 0; i < a.length; ++i) {
      switch (a[i]) {
        case "--docs" -> docs = Long.parseLong(a[++i]);
        case "--vocab" -> vocab = Integer.parseInt(a[++i]);
        case "--queries" -> nq = Integer.parseInt(a[++i]);
        case "--topk" -> topk = Integer.parseInt(a[++i]);
        case "--threshold" -> threshold = Integer.parseInt(a[++i]);
        case "--threads" -> threads = Integer.parseInt(a[++i]);
        case "--repeat" -> repeat = Integer.parseInt(a[++i]);
        case "--index" -> index = a[++i];
        case "--check" -> check = true;
        default -> throw new IllegalArgumentException(a[i]);
      }
    }
    FSDirectory dir = FSDirectory.open(Paths.get(index));
    if (!DirectoryReader.indexExists(dir)) {
      Zipf zipf = new Zipf(vocab, 1.0);
      Poisson pois = new Poisson(56.0);
      IndexWriterConfig cfg = new IndexWriterConfig(new WhitespaceAnalyzer());
      cfg.setMergePolicy(new LogByteSizeMergePolicy()); // doc order preserved across merges
      cfg.setRAMBufferSizeMB(2048);
      try (IndexWriter w = new IndexWriter(dir, cfg)) {
        StringBuilder sb = new StringBuilder();
        for (long d = 0; d < docs; ++d) {
          int len = Math.min(1000, 8 + pois.sample(u01(SEED_CORPUS ^ 0xD0C1E57L, d)));
          sb.setLength(0);
          for (int j = 0; j < len; ++j) sb.append('t').append(zipf.sample(u01(SEED_CORPUS, d * 1024L + j))).append(' ');
          Document doc = new Document();
          doc.add(new TextField("text", sb.toString(), Field.Store.NO)); // norms on, BM25Similarity default (TextFieldDef.java:134)
          w.addDocument(doc);
        }
      }
    }
    // queries: 3 distinct term ranks, log-uniform in [10, 10^4) (bench.py: synth_query_terms(nq, 3, vocab))
    Query[] queries = new Query[nq];
    for (int q = 0; q < nq; ++q) {
      int[] t = new int[3];
      int got = 0;
      long ctr = (long) q * 64L;
      while (got < 3) {
        int r = (int) Math.floor(Math.pow(10.0, 1.0 + (4.0 - 1.0) * u01(SEED_QUERIES, ctr++)));
        if (r >= vocab) r = vocab - 1;
        boolean dup = false;
        for (int j = 0; j < got; ++j) dup |= t[j] == r;
        if (!dup) t[got++] = r;
      }
      BooleanQuery.Builder b = new BooleanQuery.Builder();
      for (int x : t) b.add(new TermQuery(new Term("text", "t" + x)), BooleanClause.Occur.SHOULD);
      queries[q] = b.build();
    }
    ExecutorService searchPool = Executors.newFixedThreadPool(threads); // nrtsearch's SEARCH executor (concurrent segments)
    ExecutorService requestPool = Executors.newFixedThreadPool(threads); // nrtsearch's gRPC SERVER executor (one thread per request)
    try (DirectoryReader reader = DirectoryReader.open(dir)) {
      IndexSearcher searcher = new IndexSearcher(reader, searchPool);
      final int k = topk, thr = threshold;
      double best = 0;
      for (int rep = 0; rep <= repeat; ++rep) { // rep 0 = warm-up
        long t0 = System.nanoTime();
        List<Future<TopDocs>> fs = new ArrayList<>();
        for (Query q : queries) fs.add(requestPool.submit(() -> searcher.search(q, new TopScoreDocCollectorManager(k, null, thr))));
        TopDocs first = null;
        for (Future<TopDocs> f : fs) {
          TopDocs td = f.get();
          if (first == null) first = td;
        }
        double sec = (System.nanoTime() - t0) * 1e-9;
        if (rep > 0) best = Math.max(best, nq / sec);
        if (check && rep == 0) System.out.println("query 0: " + first.totalHits + " " + Arrays.toString(Arrays.copyOf(first.scoreDocs, Math.min(5, first.scoreDocs.length))));
      }
      System.out.printf(
          "{\"impl\": \"lucene\", \"metric\": \"BM25 queries/sec\", \"value\": %.1f, \"unit\": \"queries/s\", \"cores\": %d, \"docs\": %d, \"batch\": %d, \"top_k\": %d, \"total_hits_threshold\": %d, \"segments\": %d}%n",
          best, threads, docs, nq, topk, threshold, reader.leaves().size());
    } finally {
      searchPool.shutdown();
      requestPool.shutdown();
    }
  }
}
