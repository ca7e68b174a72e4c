"""Parity proper: CUDA path (through the C ABI) vs the CPU oracle on identical seeded inputs.
Bit-exact doc ids, bit-exact BM25 scores, exact totalHits (SURVEY.md 8c parity spec)."""
import numpy as np
import pytest

import oracle
from helpers import assert_same_hits, shard_from_token_docs
from nrtsearch_b200 import index as ix
from nrtsearch_b200.search import (BooleanQuery, BoostQuery, GpuIndex, GpuIndexSearcher, MatchAllDocsQuery, Occur,
                                   RangeQuery, RelevanceCollector, ScoreDoc, TermQuery, boolean_query_from_proto,
                                   compile_queries)

pytestmark = pytest.mark.gpu
INT_MAX = 2**31 - 1


def run_both(gpu_ctx, sh, queries, top_k, threshold=INT_MAX, search_after=None, gix=None):
    own = gix is None
    if own:
        gix = GpuIndex(gpu_ctx, sh)
    try:
        res = GpuIndexSearcher(gix).search_batch(queries, RelevanceCollector(top_k, threshold), search_after=search_after)
    finally:
        if own:
            gix.close()
    carr, ncl, qarr, nq = compile_queries(queries, search_after)
    want = oracle.search_compiled(oracle.OracleIndex(sh), carr, ncl, qarr, nq, top_k)
    got = (res.docs, res.scores, res.counts, res.total_hits, res.relation)
    return got, want


@pytest.fixture(scope="module")
def corpus():
    sh = ix.synth_text_shard(200_000, 20_000)
    sh.columns = [ix.synth_int_column(sh.n_docs)]
    sh.column_has = [None]
    return sh


def disj(terms):
    q = BooleanQuery()
    for t in terms:
        q.add(TermQuery(int(t)), Occur.SHOULD)
    return q


def test_kat_corpora_on_gpu(gpu_ctx):
    # the reference's own known answers, now through the CUDA path
    docs = ["Document1 with none of filter terms", "Document2 with term1 filter term",
            "Document1 with term2 filter term", "Document2 with both term1 and term2 filter terms"]
    sh, vocab = shard_from_token_docs([[d.lower().split() for d in docs]])
    gix = GpuIndex(gpu_ctx, sh)
    td = GpuIndexSearcher(gix).search(TermQuery(vocab[(0, "document2")]), RelevanceCollector(10, INT_MAX))
    gix.close()
    assert [sd.doc for sd in td.score_docs] == [1, 3] and td.total_hits.value == 2
    assert td.score_docs[0].score == 0.33812057971954346 and td.score_docs[1].score == 0.27725890278816223
    sh, vocab = shard_from_token_docs([["first vendor".split(), "second vendor".split()]])
    gix = GpuIndex(gpu_ctx, sh)
    td = GpuIndexSearcher(gix).search(disj([vocab[(0, "first")], vocab[(0, "vendor")]]), RelevanceCollector(100, 1000))
    gix.close()
    assert [sd.doc for sd in td.score_docs] == [0, 1]
    assert np.float32(td.score_docs[0].score) == np.float32(0.3979403)


def test_disjunction_3term_top100(gpu_ctx, corpus):
    terms = ix.synth_query_terms(256, 3, 20_000, log10_lo=0.3, log10_hi=3.5)
    got, want = run_both(gpu_ctx, corpus, [disj(t) for t in terms], 100)
    assert_same_hits(got, want, what="disjunction")
    assert (got[2] == 100).all()


def test_conjunction_with_range_filter(gpu_ctx, corpus):
    terms = ix.synth_query_terms(128, 2, 20_000, seed=ix.SEED_QUERIES + 1, log10_lo=0.3, log10_hi=2.5)
    los = (ix.synth_uniform(128, ix.SEED_RANGE) * 900_000).astype(np.int64)
    qs = [BooleanQuery().add(TermQuery(int(t[0])), Occur.MUST).add(TermQuery(int(t[1])), Occur.MUST)
          .add(RangeQuery(0, int(lo), int(lo) + 100_000), Occur.FILTER) for t, lo in zip(terms, los)]
    got, want = run_both(gpu_ctx, corpus, qs, 100)
    assert_same_hits(got, want, what="conjunction+range")
    assert got[3].sum() > 0


def test_mixed_occurs_and_boosts(gpu_ctx, corpus):
    terms = ix.synth_query_terms(64, 4, 20_000, seed=77, log10_lo=0.3, log10_hi=2.5)
    qs = []
    for i, t in enumerate(terms):
        q = BooleanQuery(minimum_number_should_match=i % 3 if i % 4 == 1 else 0)
        q.add(BoostQuery(TermQuery(int(t[0])), 1.5), Occur.MUST if i % 2 == 0 else Occur.SHOULD)
        q.add(TermQuery(int(t[1])), Occur.SHOULD)
        q.add(BoostQuery(TermQuery(int(t[2])), 0.25), Occur.SHOULD)
        q.add(TermQuery(int(t[3])), Occur.MUST_NOT if i % 3 == 0 else Occur.FILTER if i % 3 == 1 else Occur.SHOULD)
        qs.append(BoostQuery(q, 2.5) if i % 5 == 0 else q)
    got, want = run_both(gpu_ctx, corpus, qs, 50)
    assert_same_hits(got, want, what="mixed")


def test_wide_queries_use_8_slots(gpu_ctx, corpus):
    terms = ix.synth_query_terms(32, 7, 20_000, seed=78, log10_lo=0.3, log10_hi=3.0)
    qs = [disj(t) for t in terms]
    qs[3] = BooleanQuery().add(TermQuery(int(terms[3][0])), Occur.MUST)
    for t in terms[3][1:]:
        qs[3].add(TermQuery(int(t)), Occur.SHOULD)
    got, want = run_both(gpu_ctx, corpus, qs, 100)
    assert_same_hits(got, want, what="7-term")


def test_dense_drivers_matchall_range_mustnot(gpu_ctx, corpus):
    t = ix.synth_query_terms(8, 2, 20_000, seed=5, log10_lo=0.3, log10_hi=1.5)
    qs = [
        boolean_query_from_proto([]),                                                   # empty -> MatchAll MUST (score 1)
        boolean_query_from_proto([(TermQuery(int(t[0][0])), Occur.MUST_NOT)]),          # all MUST_NOT -> +MatchAll FILTER
        RangeQuery(0, 10, 5_000),                                                       # bare range: constant score 1
        BooleanQuery().add(RangeQuery(0, 0, 300_000), Occur.FILTER).add(TermQuery(int(t[1][0])), Occur.SHOULD),
        BooleanQuery().add(MatchAllDocsQuery(), Occur.SHOULD).add(TermQuery(int(t[2][0])), Occur.SHOULD),
        BooleanQuery().add(RangeQuery(0, 0, 200_000), Occur.MUST_NOT).add(TermQuery(int(t[3][0])), Occur.MUST),
        BoostQuery(MatchAllDocsQuery(), 3.0),
    ]
    got, want = run_both(gpu_ctx, corpus, qs, 20)
    assert_same_hits(got, want, what="dense")


def test_empty_and_degenerate_queries(gpu_ctx, corpus):
    import copy
    corpus = copy.copy(corpus)
    corpus.term_off = np.append(corpus.term_off, corpus.term_off[-1])   # one extra term without postings
    empty_term = corpus.n_terms - 1
    rare = empty_term - 1
    assert corpus.df(empty_term) == 0 and corpus.df(rare) > 0
    qs = [
        BooleanQuery(),                                             # no clauses at all: matches nothing
        BooleanQuery().add(TermQuery(rare), Occur.MUST_NOT),        # only MUST_NOT (no MatchAll added): nothing
        disj([empty_term]),                                         # term without postings
        disj([rare]),                                               # fewer hits than top_k
        BooleanQuery(minimum_number_should_match=3).add(TermQuery(5), Occur.SHOULD).add(TermQuery(6), Occur.SHOULD),
        BooleanQuery().add(TermQuery(3), Occur.MUST).add(TermQuery(empty_term), Occur.MUST),
    ]
    got, want = run_both(gpu_ctx, corpus, qs, 10)
    assert_same_hits(got, want, what="degenerate")
    assert list(got[2][:3]) == [0, 0, 0]


def test_search_after_paging(gpu_ctx, corpus):
    terms = ix.synth_query_terms(16, 3, 20_000, seed=9, log10_lo=0.3, log10_hi=2.0)
    qs = [disj(t) for t in terms]
    gix = GpuIndex(gpu_ctx, corpus)
    page1, want1 = run_both(gpu_ctx, corpus, qs, 20, gix=gix)
    assert_same_hits(page1, want1, what="page1")
    after = [ScoreDoc(int(page1[0][q, 19]), float(page1[1][q, 19])) for q in range(16)]
    page2, want2 = run_both(gpu_ctx, corpus, qs, 20, search_after=after, gix=gix)
    assert_same_hits(page2, want2, what="page2")
    full, _ = run_both(gpu_ctx, corpus, qs, 40, gix=gix)
    gix.close()
    assert np.array_equal(np.concatenate([page1[0], page2[0]], axis=1), full[0])   # no overlap, no gap
    assert np.array_equal(page1[3], page2[3])                                     # totalHits unchanged by paging


def test_deleted_docs_and_ties(gpu_ctx):
    # many identical docs -> exact score ties; liveDocs removes hits but not statistics
    docs = [["x", "y"] if i % 3 else ["x", "z"] for i in range(5000)]
    live = np.ones(5000, np.uint8)
    live[::7] = 0
    sh, vocab = shard_from_token_docs([docs], live_docs=live)
    qs = [disj([vocab[(0, "x")], vocab[(0, "y")]]), TermQuery(vocab[(0, "x")]), disj([vocab[(0, "z")]])]
    got, want = run_both(gpu_ctx, sh, qs, 64)
    assert_same_hits(got, want, what="ties")


def test_multi_field_omit_norms_and_tf_saturation(gpu_ctx):
    rng = np.random.default_rng(3)
    body, title = [], []
    for d in range(3000):
        n = int(rng.integers(3, 60))
        toks = [f"w{int(x)}" for x in rng.zipf(1.3, n) if x < 200]
        if d % 500 == 0:
            toks += ["w1"] * (300 + d // 10)      # tf >= 255: byte saturates, exception list path
        body.append(toks or ["w1"])
        title.append([f"w{int(x)}" for x in rng.zipf(1.5, 4) if x < 50] if d % 3 else [])
    sh, vocab = shard_from_token_docs([body, title], omit_norms=[False, True])
    t = lambda f, w: TermQuery(vocab[(f, w)])
    qs = [disj([vocab[(0, "w1")], vocab[(1, "w1")], vocab[(0, "w2")]]),
          BooleanQuery().add(t(0, "w1"), Occur.MUST).add(t(1, "w2"), Occur.SHOULD),
          BooleanQuery().add(t(1, "w1"), Occur.MUST).add(t(0, "w3"), Occur.MUST),
          t(0, "w1")]
    got, want = run_both(gpu_ctx, sh, qs, 100)
    assert_same_hits(got, want, what="multi-field")


def test_multi_slice_index(gpu_ctx):
    # > 1,048,576 docs => several doc slices per query + slice merge
    sh = ix.synth_text_shard(2_300_000, 50_000, min_len=4, poisson_mean=12.0)
    terms = ix.synth_query_terms(64, 3, 50_000, seed=11, log10_lo=0.3, log10_hi=3.5)
    got, want = run_both(gpu_ctx, sh, [disj(t) for t in terms], 100)
    assert_same_hits(got, want, what="multi-slice")


def test_argument_errors(gpu_ctx, corpus):
    from nrtsearch_b200 import NrtGpuError, NrtGpuUnsupported
    gix = GpuIndex(gpu_ctx, corpus)
    s = GpuIndexSearcher(gix)
    with pytest.raises(NrtGpuError, match="numHits must be > 0"):
        s.search_batch([TermQuery(1)], RelevanceCollector(0))
    with pytest.raises(NrtGpuError, match="term id out of range"):
        s.search_batch([TermQuery(10**8)], RelevanceCollector(10))
    with pytest.raises(NrtGpuUnsupported):
        s.search_batch([BooleanQuery().add(BooleanQuery(), Occur.MUST)], RelevanceCollector(10))
    with pytest.raises(NrtGpuUnsupported):
        s.search_batch([disj(range(1, 11))], RelevanceCollector(10))
    gix.close()


def test_top_scores_mode_skips_lists_but_keeps_topk(gpu_ctx):
    """totalHitsThreshold < MAX (reference default 1000): MAXSCORE list skipping may run; the (doc, score) lists must
    still equal the exhaustive oracle, totalHits must be exact when EQUAL_TO and a lower bound > threshold otherwise
    (TotalHitsThresholdTest.java:72-100 semantics)."""
    sh = ix.synth_text_shard(2_300_000, 50_000, min_len=4, poisson_mean=12.0)   # 3 slices: later slices see a warm theta
    terms = ix.synth_query_terms(96, 3, 50_000, seed=21, log10_lo=0.3, log10_hi=4.0)
    qs = [disj(t) for t in terms]
    gix = GpuIndex(gpu_ctx, sh)
    batch = GpuIndexSearcher(gix).prepare(qs, RelevanceCollector(100, 1000))
    batch.run()
    res = batch.fetch()
    batch.close()
    gix.close()
    carr, ncl, qarr, nq = compile_queries(qs)
    want = oracle.search_compiled(oracle.OracleIndex(sh), carr, ncl, qarr, nq, 100)
    got = (res.docs, res.scores, res.counts, res.total_hits, res.relation)
    assert_same_hits(got, want, check_total=False, what="TOP_SCORES")
    eq = res.relation == 0
    assert np.array_equal(res.total_hits[eq], want[3][eq])
    assert (res.total_hits[~eq] <= want[3][~eq]).all() and (res.total_hits[~eq] > 1000).all()
    assert (~eq).any(), "expected at least one query to skip a non-essential list"


def test_top_scores_mixed_term_counts_and_search_after(gpu_ctx):
    """TOP_SCORES mode over 1..4-term disjunctions of dense (plane-served), medium and rare terms: window mode, tf-plane
    mode and sparse (binary-search merge) mode all appear inside one batch; paging with searchAfter on top. The
    (doc, score) lists must equal the exhaustive oracle bit for bit."""
    sh = ix.synth_text_shard(1_300_000, 30_000, min_len=4, poisson_mean=14.0)   # 3 slices of 524,288 docs
    rng = np.random.default_rng(5)
    qs = []
    for n_terms in (1, 2, 3, 4):
        for _ in range(96):   # 386 queries x 3 slices > the 296 CTAs resident at once: later work items see a warm theta
            ranks = np.unique(np.floor(10 ** rng.uniform(0.0, 4.2, size=n_terms)).astype(np.int64).clip(1, 29_999))
            qs.append(disj(ranks))
    qs.append(disj([1, 2, 3]))          # three very dense lists
    qs.append(disj([1, 20_000, 25_000]))  # one dense + two rare lists
    gix = GpuIndex(gpu_ctx, sh)
    s = GpuIndexSearcher(gix)
    res = s.search_batch(qs, RelevanceCollector(50, 200))
    carr, ncl, qarr, nq = compile_queries(qs)
    want = oracle.search_compiled(oracle.OracleIndex(sh), carr, ncl, qarr, nq, 50)
    got = (res.docs, res.scores, res.counts, res.total_hits, res.relation)
    assert_same_hits(got, want, check_total=False, what="TOP_SCORES mixed")
    assert (res.relation != 0).any()
    # second page after the 50th hit of every query that has one
    after = [ScoreDoc(int(res.docs[q, res.counts[q] - 1]), float(res.scores[q, res.counts[q] - 1])) if res.counts[q] == 50 else None
             for q in range(len(qs))]
    sel = [q for q in range(len(qs)) if after[q] is not None]
    res2 = s.search_batch([qs[q] for q in sel], RelevanceCollector(50, 200), search_after=[after[q] for q in sel])
    gix.close()
    carr, ncl, qarr, nq = compile_queries([qs[q] for q in sel], [after[q] for q in sel])
    want2 = oracle.search_compiled(oracle.OracleIndex(sh), carr, ncl, qarr, nq, 50)
    got2 = (res2.docs, res2.scores, res2.counts, res2.total_hits, res2.relation)
    assert_same_hits(got2, want2, check_total=False, what="TOP_SCORES mixed page 2")
