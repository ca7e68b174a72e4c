// Deterministic synthetic corpus / query / vector generators.
//
// Host-only C ABI, OpenMP. Every value is a pure function of (seed, index), so the
// output does not depend on the thread count. Used by bench.py and tests/ to make
// the inputs that BASELINE.json's configs name; NOT part of the search hot path.
//
// SmallFloat.intToByte4 is restated here only to produce the norms column exactly
// the way an index writer would (Lucene BM25Similarity.computeNorm; norms enabled
// by src/main/java/com/yelp/nrtsearch/server/field/TextFieldDef.java:134).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// counter-based stream: u in [0,1) with 53 bits
inline uint64_t rnd(uint64_t seed, uint64_t counter) { return mix64(seed ^ mix64(counter)); }
inline double u01(uint64_t seed, uint64_t counter) {
  return (double)(rnd(seed, counter) >> 11) * (1.0 / 9007199254740992.0);
}

inline int bitlen64(uint64_t x) { return x ? 64 - __builtin_clzll(x) : 0; }
inline int long_to_int4(uint64_t i) {
  int nbits = bitlen64(i);
  if (nbits < 4) return (int)i;
  int shift = nbits - 4;
  int enc = (int)(i >> shift) & 0x07;
  return enc | ((shift + 1) << 3);
}
inline uint8_t int_to_byte4(int i) {
  const int NUM_FREE = 255 - long_to_int4(0x7fffffff);  // 24
  if (i < NUM_FREE) return (uint8_t)i;
  return (uint8_t)(NUM_FREE + long_to_int4((uint64_t)(i - NUM_FREE)));
}

struct Zipf {
  std::vector<double> cdf;     // cdf[r] = P(rank <= r), r = 0..V-1
  std::vector<uint32_t> guide; // guide[b] = lower_bound(cdf, b / G)
  int V;
  static constexpr int GBITS = 20;
  explicit Zipf(int V_, double s) : cdf(V_), guide((1u << GBITS) + 1), V(V_) {
    double h = 0;
    for (int r = 0; r < V; ++r) { h += 1.0 / std::pow((double)(r + 1), s); cdf[r] = h; }
    for (int r = 0; r < V; ++r) cdf[r] /= h;
    cdf[V - 1] = 1.0;
    uint32_t r = 0;
    const uint32_t G = 1u << GBITS;
    for (uint32_t b = 0; b <= G; ++b) {
      double x = (double)b / (double)G;
      while (r < (uint32_t)V - 1 && cdf[r] < x) ++r;
      guide[b] = r;
    }
  }
  inline int sample(double u) const {
    uint32_t b = (uint32_t)(u * (double)(1u << GBITS));
    uint32_t lo = guide[b], hi = guide[b + 1];
    // first r in [lo, hi] with cdf[r] > u  (u < cdf[hi] is guaranteed unless hi==V-1)
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    return (int)lo;
  }
};

struct Poisson {
  std::vector<double> cdf;
  explicit Poisson(double lam) {
    double p = std::exp(-lam), c = p;
    cdf.push_back(c);
    for (int k = 1; k < 1000; ++k) { p *= lam / k; c += p; cdf.push_back(c); if (1.0 - c < 1e-17 && k > lam) break; }
  }
  inline int sample(double u) const {
    return (int)(std::lower_bound(cdf.begin(), cdf.end(), u, [](double c, double x) { return c <= x; }) - cdf.begin());
  }
};

struct Corpus {
  int64_t n_docs; int64_t doc_begin; int vocab; uint64_t seed; int min_len; double pois_mean;
  Zipf zipf; Poisson pois;
  Corpus(int64_t n, int64_t d0, int v, uint64_t s, int ml, double pm, double zs)
      : n_docs(n), doc_begin(d0), vocab(v), seed(s), min_len(ml), pois_mean(pm), zipf(v, zs), pois(pm) {}
  inline int doc_len(int64_t d) const {
    int L = min_len + pois.sample(u01(seed ^ 0xD0C1E57ull, (uint64_t)d));
    return L > 1000 ? 1000 : L;
  }
  // fills terms[] sorted; returns length
  inline int doc_tokens(int64_t d, int* terms) const {
    int L = doc_len(d);
    for (int j = 0; j < L; ++j) terms[j] = zipf.sample(u01(seed, (uint64_t)d * 1024ull + (uint64_t)j));
    std::sort(terms, terms + L);
    return L;
  }
};

}  // namespace

extern "C" {

struct nrtsynth_corpus {
  Corpus* c;
  int n_chunks;
  std::vector<std::vector<uint32_t>>* chunk_counts;  // per chunk: per-term posting counts
};

// Pass 1: count. Returns handle; fills df[vocab] (int64), norms[n_docs], *sum_ttf, *n_postings.
// The shard holds global docs [doc_begin, doc_begin + n_docs); stored doc ids are shard-local.
nrtsynth_corpus* nrtsynth_corpus_begin(int64_t n_docs, int64_t doc_begin, int vocab, uint64_t seed, int min_len,
                                       double poisson_mean, double zipf_s, int64_t* df /*[vocab]*/,
                                       uint8_t* norms /*[n_docs]*/, int64_t* sum_ttf,
                                       int64_t* n_postings) {
  auto* h = new nrtsynth_corpus;
  h->c = new Corpus(n_docs, doc_begin, vocab, seed, min_len, poisson_mean, zipf_s);
  int nth = 1;
#ifdef _OPENMP
  nth = omp_get_max_threads();
#endif
  int n_chunks = (int)std::min<int64_t>(std::max<int64_t>(1, n_docs / 4096), (int64_t)nth);
  h->n_chunks = n_chunks;
  h->chunk_counts = new std::vector<std::vector<uint32_t>>(n_chunks, std::vector<uint32_t>(vocab, 0));
  std::vector<int64_t> ttf(n_chunks, 0);
#pragma omp parallel for schedule(static, 1)
  for (int ch = 0; ch < n_chunks; ++ch) {
    int64_t d0 = n_docs * ch / n_chunks, d1 = n_docs * (ch + 1) / n_chunks;
    auto& cnt = (*h->chunk_counts)[ch];
    int terms[1024];
    int64_t t = 0;
    for (int64_t d = d0; d < d1; ++d) {
      int L = h->c->doc_tokens(d + doc_begin, terms);
      norms[d] = int_to_byte4(L);
      t += L;
      for (int j = 0; j < L;) { int k = j + 1; while (k < L && terms[k] == terms[j]) ++k; cnt[terms[j]]++; j = k; }
    }
    ttf[ch] = t;
  }
  int64_t P = 0, T = 0;
  for (int t = 0; t < vocab; ++t) {
    int64_t s = 0;
    for (int ch = 0; ch < n_chunks; ++ch) s += (*h->chunk_counts)[ch][t];
    df[t] = s; P += s;
  }
  for (int ch = 0; ch < n_chunks; ++ch) T += ttf[ch];
  *sum_ttf = T; *n_postings = P;
  return h;
}

// Pass 2: fill CSR postings. term_off[vocab+1] is written; docs/freqs have n_postings entries.
void nrtsynth_corpus_fill(nrtsynth_corpus* h, int64_t* term_off, int32_t* docs, int32_t* freqs) {
  Corpus* c = h->c;
  int V = c->vocab, n_chunks = h->n_chunks;
  term_off[0] = 0;
  for (int t = 0; t < V; ++t) {
    int64_t s = 0;
    for (int ch = 0; ch < n_chunks; ++ch) s += (*h->chunk_counts)[ch][t];
    term_off[t + 1] = term_off[t] + s;
  }
  // per-chunk write cursors (reuse counts storage as 64-bit cursors)
  std::vector<std::vector<int64_t>> cur(n_chunks, std::vector<int64_t>());
  for (int ch = 0; ch < n_chunks; ++ch) cur[ch].resize(V);
  for (int t = 0; t < V; ++t) {
    int64_t o = term_off[t];
    for (int ch = 0; ch < n_chunks; ++ch) { cur[ch][t] = o; o += (*h->chunk_counts)[ch][t]; }
  }
#pragma omp parallel for schedule(static, 1)
  for (int ch = 0; ch < n_chunks; ++ch) {
    int64_t d0 = c->n_docs * ch / n_chunks, d1 = c->n_docs * (ch + 1) / n_chunks;
    auto& w = cur[ch];
    int terms[1024];
    for (int64_t d = d0; d < d1; ++d) {
      int L = c->doc_tokens(d + c->doc_begin, terms);
      for (int j = 0; j < L;) {
        int k = j + 1; while (k < L && terms[k] == terms[j]) ++k;
        int64_t p = w[terms[j]]++;
        docs[p] = (int32_t)d; freqs[p] = k - j;
        j = k;
      }
    }
  }
}

void nrtsynth_corpus_end(nrtsynth_corpus* h) {
  if (!h) return;
  delete h->chunk_counts; delete h->c; delete h;
}

// int32 doc-value column: value = floor(range * u(seed, global doc))
void nrtsynth_int_column(int64_t n_docs, int64_t doc_begin, uint64_t seed, int32_t range, int32_t* out) {
#pragma omp parallel for schedule(static)
  for (int64_t d = 0; d < n_docs; ++d) out[d] = (int32_t)((double)range * u01(seed, (uint64_t)(d + doc_begin)));
}

// Queries: nq x terms_per_query distinct term ids, rank = floor(10^(lo + (hi-lo)*u)) clamped to vocab-1.
void nrtsynth_queries(int nq, int terms_per_query, uint64_t seed, double log10_lo, double log10_hi,
                      int vocab, int32_t* out /*[nq*terms_per_query]*/) {
  for (int q = 0; q < nq; ++q) {
    int got = 0; uint64_t ctr = (uint64_t)q * 64ull;
    while (got < terms_per_query) {
      double u = u01(seed, ctr++);
      int r = (int)std::floor(std::pow(10.0, log10_lo + (log10_hi - log10_lo) * u));
      if (r >= vocab) r = vocab - 1;
      bool dup = false;
      for (int j = 0; j < got; ++j) dup |= (out[q * terms_per_query + j] == r);
      if (!dup) out[q * terms_per_query + got++] = r;
    }
  }
}

// uniform doubles in [0,1): out[i] = u(seed, i)
void nrtsynth_uniform(int64_t n, uint64_t seed, double* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) out[i] = u01(seed, (uint64_t)i);
}

// fp32 i.i.d. N(0,1) (Box-Muller on two successive counters), row-major [n, dims];
// value_begin (even) = global index of out[0], so shards of one matrix can be generated independently
void nrtsynth_normal_f32(int64_t n_values, int64_t value_begin, uint64_t seed, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_values; i += 2) {
    double u1 = u01(seed, (uint64_t)(i + value_begin)), u2 = u01(seed, (uint64_t)(i + value_begin) + 1);
    if (u1 < 1e-300) u1 = 1e-300;
    double r = std::sqrt(-2.0 * std::log(u1)), a = 6.283185307179586476925 * u2;
    out[i] = (float)(r * std::cos(a));
    if (i + 1 < n_values) out[i + 1] = (float)(r * std::sin(a));
  }
}

}  // extern "C"
