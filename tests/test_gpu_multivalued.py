"""Multi-valued numeric doc values (SORTED_NUMERIC; reference NumberFieldDef multiValued fields): a range clause matches a
doc when ANY of its values lies in [lo, hi] (Lucene SortedNumericDocValuesRangeQuery, reached through
IntFieldDef.getRangeQuery :124-158). GPU top-k / totals against the oracle; sort / aggregations / fetch on such a column
answer UNSUPPORTED; nrtgpu_index_build rejects malformed shards instead of reading out of bounds."""
import numpy as np
import pytest

import oracle
from helpers import assert_same_hits
from nrtsearch_b200 import _native as N
from nrtsearch_b200 import index as ix
from nrtsearch_b200.search import (BooleanQuery, GpuIndex, GpuIndexSearcher, MatchAllDocsQuery, Occur, RangeQuery, RelevanceCollector,
                                   SortFieldCollector, SortType, TermQuery, compile_queries)

pytestmark = pytest.mark.gpu
INT_MAX = 2**31 - 1


def multi_shard(n_docs=300_000, vocab=8000, seed=5):
    sh = ix.synth_text_shard(n_docs, vocab)
    rng = np.random.default_rng(seed)
    cnt = rng.integers(0, 5, n_docs)                 # 0..4 values per doc (0 = the doc has none)
    off = np.zeros(n_docs + 1, np.int64)
    np.cumsum(cnt, out=off[1:])
    vals = rng.integers(-1000, 1000, int(off[-1])).astype(np.int64)
    vals[::97] += 1 << 40                            # a few values outside int32
    doc_of = np.repeat(np.arange(n_docs), cnt)
    order = np.lexsort((vals, doc_of))               # ascending within a doc
    vals = vals[order]
    single = rng.integers(0, 100, n_docs).astype(np.int64)
    sh.columns = [vals, single]
    sh.column_has = [None, None]
    sh.column_offsets = [off, None]
    return sh


def build_queries(vocab, n=90, seed=3):
    terms = ix.synth_query_terms(n, 2, vocab, seed=seed, log10_lo=0.3, log10_hi=3.3)
    rng = np.random.default_rng(seed)
    qs = []
    for i, t in enumerate(terms):
        q = BooleanQuery()
        lo = int(rng.integers(-1100, 900))
        hi = lo + int(rng.integers(0, 400))
        if i % 10 == 0:
            lo, hi = (1 << 40) - 1000, (1 << 40) + 1000
        if i % 4 == 3:                                # range-led: no scoring term at all
            q.add(MatchAllDocsQuery(), Occur.MUST)
            q.add(RangeQuery(0, lo, hi), Occur.FILTER)
        else:
            q.add(TermQuery(int(t[0])), Occur.MUST if i % 2 else Occur.SHOULD)
            q.add(TermQuery(int(t[1])), Occur.SHOULD)
            q.add(RangeQuery(0, lo, hi), Occur.MUST_NOT if i % 4 == 2 else Occur.FILTER)
        if i % 5 == 0:
            q.add(RangeQuery(1, 10, 60), Occur.FILTER)
        qs.append(q)
    return qs


def test_any_value_in_range_matches(gpu_ctx):
    sh = multi_shard()
    qs = build_queries(8000)
    carr, ncl, qarr, nq = compile_queries(qs)
    want = oracle.search_compiled(oracle.OracleIndex(sh), carr, ncl, qarr, nq, 20)
    # the oracle's linear scan against plain numpy on a few range-only queries
    off, vals = sh.column_offsets[0], sh.columns[0]
    for q in (3, 7, 11):
        c = [x.query for x in qs[q].clauses if isinstance(x.query, RangeQuery) and x.query.column == 0][0]
        hit = np.add.reduceat(((vals >= c.lower) & (vals <= c.upper)).astype(np.int64), np.minimum(off[:-1], len(vals) - 1)) * (np.diff(off) > 0)
        assert want[3][q] == int((hit > 0).sum())
    gi = GpuIndex(gpu_ctx, sh)
    s = GpuIndexSearcher(gi)
    for thr in (INT_MAX, 100):
        res = s.search_batch(qs, RelevanceCollector(20, thr))
        assert_same_hits((res.docs, res.scores, res.counts, res.total_hits, res.relation), want, check_total=(thr == INT_MAX), what=f"multi-valued thr={thr}")
    # two doc-range sub-shards merge to the same answer (offsets re-based per shard)
    parts = [GpuIndex(gpu_ctx, sh.doc_range(0, 130_000)), GpuIndex(gpu_ctx, sh.doc_range(130_000, 300_000))]
    from nrtsearch_b200.search import GpuLeafSearcher
    ls = GpuLeafSearcher(gpu_ctx, parts)
    res = ls.search_batch(qs, RelevanceCollector(20, INT_MAX))
    assert_same_hits((res.docs, res.scores, res.counts, res.total_hits, res.relation), want, what="multi-valued, 2 leaves")
    # collectors that need ONE value per doc refuse the column
    with pytest.raises(N.NrtGpuError) as e:
        s.search_sorted(qs[:4], SortFieldCollector(10, SortType(0)))
    assert e.value.status == 3
    with pytest.raises(N.NrtGpuError) as e:
        s.fetch_columns([0], np.array([1, 2, 3], np.int32))
    assert e.value.status == 3


def test_index_build_rejects_malformed_shards(gpu_ctx):
    def expect_invalid(mutate):
        sh = ix.synth_text_shard(5000, 300)
        mutate(sh)
        with pytest.raises(N.NrtGpuError) as e:
            GpuIndex(gpu_ctx, sh)
        assert e.value.status == 1

    def unsorted(sh):
        a = int(sh.term_off[5])
        sh.post_docs = sh.post_docs.copy(); sh.post_docs[a], sh.post_docs[a + 1] = sh.post_docs[a + 1], sh.post_docs[a]
    def out_of_range(sh):
        sh.post_docs = sh.post_docs.copy(); sh.post_docs[int(sh.term_off[9]) - 1] = sh.n_docs
    def bad_off0(sh):
        sh.term_off = sh.term_off.copy(); sh.term_off[0] = 1
    def too_many_vectors(sh):
        sh.vectors = np.zeros((sh.n_docs + 1, 8), np.float32)
    def vec_docs_range(sh):
        sh.vectors = np.ones((4, 8), np.float32); sh.vec_docs = np.array([0, 1, 2, sh.n_docs], np.int32)
    def mv_descending(sh):
        sh.columns = [np.array([3, 1] + [0] * (sh.n_docs - 1), np.int64)]
        sh.column_has = [None]
        sh.column_offsets = [np.concatenate([[0, 2], np.arange(3, sh.n_docs + 2)]).astype(np.int64)]
    for m in (unsorted, out_of_range, bad_off0, too_many_vectors, vec_docs_range, mv_descending):
        expect_invalid(m)
