"""posting_probe_kernel parity: the paths the generic parity tests reach only by luck -- several runs per work item
(long lists without a tf plane that exceed the shared-memory stage), short lists too long for the stage reserve
(searched in global memory), exact totalHits in ScoreMode.COMPLETE with a dense non-essential list that is never
swept, and leap-frog conjunctions with MUST_NOT / optional clauses. Oracle = exhaustive CPU evaluation."""
import numpy as np
import pytest

import oracle
from helpers import assert_same_hits
from nrtsearch_b200 import index as ix
from nrtsearch_b200.search import (BooleanQuery, BoostQuery, GpuIndex, GpuIndexSearcher, Occur, RangeQuery,
                                   RelevanceCollector, ScoreDoc, TermQuery, compile_queries)

pytestmark = pytest.mark.gpu
INT_MAX = 2**31 - 1


@pytest.fixture(scope="module")
def corpus():
    # 300K docs: tf planes for df >= 4688, skip data (granule offsets) for df >= 4096, one 293-granule slice
    sh = ix.synth_text_shard(300_000, 8_000, min_len=6, poisson_mean=30.0)
    sh.columns = [ix.synth_int_column(sh.n_docs)]
    sh.column_has = [None]
    return sh


def pools(sh):
    df = np.diff(sh.term_off)
    n = sh.n_docs
    dense = np.nonzero(df * 64 >= n)[0]                         # tf plane
    long_ = np.nonzero((df >= 4096) & (df * 64 < n))[0]         # staged per run, granule-narrowed searches
    big_short = np.nonzero((df >= 3000) & (df < 4096))[0]       # no skip data; two of them exceed the stage reserve
    short = np.nonzero((df >= 50) & (df < 1500))[0]
    return dense, long_, big_short, short


def disj(terms):
    q = BooleanQuery()
    for t in terms:
        q.add(TermQuery(int(t)), Occur.SHOULD)
    return q


def run(gpu_ctx, sh, qs, top_k, threshold, search_after=None):
    gix = GpuIndex(gpu_ctx, sh)
    try:
        res = GpuIndexSearcher(gix).search_batch(qs, RelevanceCollector(top_k, threshold), search_after=search_after)
    finally:
        gix.close()
    carr, ncl, qarr, nq = compile_queries(qs, search_after)
    want = oracle.search_compiled(oracle.OracleIndex(sh), carr, ncl, qarr, nq, top_k)
    return (res.docs, res.scores, res.counts, res.total_hits, res.relation), want


def mixed_queries(sh, rng, n):
    dense, long_, big_short, short = pools(sh)
    assert len(dense) >= 3 and len(long_) >= 1 and len(big_short) >= 2 and len(short) >= 4, (len(dense), len(long_), len(big_short), len(short))
    qs = []
    for i in range(n):
        kind = i % 8
        pick = lambda pool, k: list(rng.choice(pool, size=min(k, len(pool)), replace=False))
        if kind == 0: terms = pick(long_, 3) if len(long_) >= 3 else pick(long_, 1) + pick(big_short, 2)   # > 8192 staged postings: runs
        elif kind == 1: terms = pick(big_short, 3) if len(big_short) >= 3 else pick(big_short, 2) + pick(short, 1)   # stage reserve overflow
        elif kind == 2: terms = pick(dense, 2) + pick(short, 1)
        elif kind == 3: terms = pick(dense, 1) + pick(long_, 1) + pick(big_short, 1) + pick(short, 1)
        elif kind == 4: terms = pick(short, 1)
        elif kind == 5: terms = pick(dense, 3)
        elif kind == 6: terms = pick(long_, 1) + pick(big_short, 2)
        else:
            t = pick(dense, 1) + pick(short, 1)
            terms = t + [t[0]]           # the same term twice: two slots over one list
        qs.append(disj(terms))
    return qs


@pytest.mark.parametrize("threshold", [INT_MAX, 50])
def test_disjunction_roles_runs_and_counts(gpu_ctx, corpus, threshold):
    rng = np.random.default_rng(17)
    qs = mixed_queries(corpus, rng, 160)
    got, want = run(gpu_ctx, corpus, qs, 20, threshold)
    assert_same_hits(got, want, check_total=False, what=f"probe disjunction thr={threshold}")
    eq = got[4] == 0
    assert np.array_equal(got[3][eq], want[3][eq])
    if threshold == INT_MAX:
        assert eq.all(), "ScoreMode.COMPLETE must report exact counts"
    else:
        assert (got[3][~eq] <= want[3][~eq]).all() and (got[3][~eq] > threshold).all()


def test_search_after_through_probe(gpu_ctx, corpus):
    rng = np.random.default_rng(18)
    qs = mixed_queries(corpus, rng, 48)
    page1, want1 = run(gpu_ctx, corpus, qs, 15, INT_MAX)
    assert_same_hits(page1, want1, what="page 1")
    sel = [q for q in range(len(qs)) if page1[2][q] == 15]
    after = [ScoreDoc(int(page1[0][q, 14]), float(page1[1][q, 14])) for q in sel]
    page2, want2 = run(gpu_ctx, corpus, [qs[q] for q in sel], 15, INT_MAX, search_after=after)
    assert_same_hits(page2, want2, what="page 2")


def test_conjunctions_leapfrog(gpu_ctx, corpus):
    rng = np.random.default_rng(19)
    dense, long_, big_short, short = pools(corpus)
    allp = np.concatenate([dense, long_, big_short, short])
    qs = []
    for i in range(120):
        t = rng.choice(allp, size=4, replace=False)
        q = BooleanQuery(minimum_number_should_match=1 if i % 7 == 3 else 0)
        q.add(TermQuery(int(t[0])), Occur.MUST)
        q.add(BoostQuery(TermQuery(int(t[1])), 1.75), Occur.MUST if i % 2 == 0 else Occur.SHOULD)
        if i % 3 == 0:
            q.add(TermQuery(int(t[2])), Occur.MUST_NOT)
        elif i % 3 == 1:
            q.add(TermQuery(int(t[2])), Occur.SHOULD)
        if i % 4 == 0:
            lo = int(rng.integers(0, 800_000))
            q.add(RangeQuery(0, lo, lo + 150_000), Occur.FILTER)
        if i % 5 == 0:
            q.add(TermQuery(int(t[3])), Occur.FILTER)
        qs.append(q)
    # pure SHOULD lists with an excluded list: every optional list leads, the MUST_NOT list is only probed
    for i in range(24):
        t = rng.choice(allp, size=3, replace=False)
        qs.append(BooleanQuery().add(TermQuery(int(t[0])), Occur.SHOULD).add(TermQuery(int(t[1])), Occur.SHOULD)
                  .add(TermQuery(int(t[2])), Occur.MUST_NOT))
    got, want = run(gpu_ctx, corpus, qs, 25, INT_MAX)
    assert_same_hits(got, want, what="probe generic")
    got, want = run(gpu_ctx, corpus, qs, 25, 100)
    assert_same_hits(got, want, what="probe generic TOP_SCORES")


def test_multi_slice_top_scores_and_complete_agree(gpu_ctx):
    sh = ix.synth_text_shard(1_200_000, 40_000, min_len=4, poisson_mean=14.0)   # 3 slices of 400K docs, warm-up items
    rng = np.random.default_rng(23)
    qs = []
    for _ in range(300):
        n_terms = int(rng.integers(1, 5))
        ranks = np.unique(np.floor(10 ** rng.uniform(0.0, 4.4, size=n_terms)).astype(np.int64).clip(1, 39_999))
        qs.append(disj(ranks))
    for threshold in (INT_MAX, 500):
        got, want = run(gpu_ctx, sh, qs, 100, threshold)
        assert_same_hits(got, want, check_total=False, what=f"3 slices thr={threshold}")
        eq = got[4] == 0
        assert np.array_equal(got[3][eq], want[3][eq])
        if threshold == INT_MAX:
            assert eq.all()


def test_pruned_total_hits_is_a_lower_bound_above_the_threshold(gpu_ctx):
    """TOP_SCORES with the sweep warm-up (threshold from the rarest list's postings, pruning from the first work item because
    the longest list already proves totalHits > totalHitsThreshold): the page equals the exhaustive oracle's, and a
    GREATER_THAN_OR_EQUAL_TO count is what the relation says -- above the threshold, never above the exact count
    (TopDocs.totalHits contract; LazyQueueTopScoreDocCollector.java:129-143)."""
    sh = ix.synth_text_shard(700_000, 20_000, min_len=6, poisson_mean=40.0)   # 2 slices; warm-up items are scheduled
    rng = np.random.default_rng(23)
    qs = []
    for i in range(160):
        n_terms = 2 + i % 3
        ranks = np.unique(np.floor(10 ** rng.uniform(0.3, 4.0, size=n_terms)).astype(np.int64).clip(1, 19_999))
        qs.append(disj(ranks))
    qs.append(disj([3, 15_000]))            # a dense list (known hits >> threshold) + a rare one (swept by the warm-up)
    qs.append(disj([18_000, 19_000]))       # two rare lists: fewer matching docs than the threshold -> exact count
    thr = 300
    got, want = run(gpu_ctx, sh, qs, 40, thr)
    assert_same_hits(got, want, check_total=False, what="sweep warm-up")
    total, rel = got[3], got[4]
    exact = want[3]                         # the oracle ran ScoreMode.COMPLETE
    gte = rel != 0
    assert gte.any() and (~gte).any()
    assert np.array_equal(total[~gte], exact[~gte]), "EQUAL_TO counts must be exact"
    assert (total[gte] > thr).all(), "a pruned search reports more hits than the threshold"
    assert (total[gte] <= exact[gte]).all(), "a GREATER_THAN_OR_EQUAL_TO count is a lower bound"
