"""Sort-by-field top-k: TopFieldCollector semantics of SortFieldCollector.java:44-105 with the
numeric SortFields of NumberFieldDef.java:266-278 -- value order (reverse or not), docs without a value sorting as the
FieldDef's missing value, ties by doc id, searchAfter on (value, doc). CUDA path vs the exhaustive oracle, bit-exact."""
import numpy as np
import pytest

import oracle
from nrtsearch_b200 import index as ix
from nrtsearch_b200.search import (BooleanQuery, FieldDoc, GpuIndex, GpuIndexSearcher, MatchAllDocsQuery, Occur, RangeQuery,
                                   SortFieldCollector, SortType, TermQuery, compile_queries, double_to_sortable_long,
                                   float_to_sortable_int)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(gpu_ctx):
    n = 300_000
    sh = ix.synth_text_shard(n, 8_000, min_len=6, poisson_mean=30.0)
    rng = np.random.default_rng(41)
    price = ix.synth_int_column(n, value_range=5_000)                      # int column, heavy ties
    rating = np.array([float_to_sortable_int(x) for x in rng.normal(0, 3, 4096)], np.int64)[rng.integers(0, 4096, n)]   # float column
    stamp = rng.integers(-2**62, 2**62, n, dtype=np.int64)                  # long column with the extremes present
    stamp[rng.integers(0, n, 50)] = -(2**63)
    stamp[rng.integers(0, n, 50)] = 2**63 - 1
    has_rating = (rng.random(n) < 0.7).astype(np.uint8)
    has_stamp = (rng.random(n) < 0.9).astype(np.uint8)
    sh.columns = [price, rating, stamp]
    sh.column_has = [None, has_rating, has_stamp]
    terms = ix.synth_query_terms(40, 3, 8_000, log10_lo=0.3, log10_hi=3.3)
    qs = []
    for i, t in enumerate(terms):
        if i % 5 == 0:
            qs.append(BooleanQuery().add(TermQuery(int(t[0])), Occur.MUST).add(RangeQuery(0, 100, 3_000), Occur.FILTER))
        elif i % 5 == 1:
            qs.append(BooleanQuery().add(TermQuery(int(t[0])), Occur.SHOULD).add(TermQuery(int(t[1])), Occur.SHOULD).add(TermQuery(int(t[2])), Occur.MUST_NOT))
        elif i % 5 == 2:
            qs.append(MatchAllDocsQuery() if i % 2 else RangeQuery(0, 0, 1_000))     # no posting list can lead
        else:
            q = BooleanQuery()
            for x in t:
                q.add(TermQuery(int(x)), Occur.SHOULD)
            qs.append(q)
    gix = GpuIndex(gpu_ctx, sh)
    yield sh, qs, gix
    gix.close()


SORTS = [SortType(0, False, False, "int"), SortType(0, True, True, "int"), SortType(1, False, True, "float"),
         SortType(1, True, False, "float"), SortType(2, False, False, "long"), SortType(2, True, True, "long"),
         SortType(2, False, True, "long"), SortType("docid", False), SortType("docid", True)]


def oracle_sorted(sh, qs, k, st, after=None):
    from nrtsearch_b200.search import ScoreDoc
    sd = None if after is None else [None if a is None else ScoreDoc(a.doc, 0.0) for a in after]
    carr, ncl, qarr, nq = compile_queries(qs, sd)
    av = None if after is None else [0 if a is None else a.value for a in after]
    docid = st.field == "docid"
    return oracle.search_sorted(oracle.OracleIndex(sh), carr, ncl, qarr, nq, k, 2 if docid else 1, 0 if docid else st.field, st.reverse,
                                0 if docid else st.missing_value(), av)


@pytest.mark.parametrize("st", SORTS, ids=lambda s: f"{s.field}-{'desc' if s.reverse else 'asc'}-{'last' if s.missing_last else 'first'}")
def test_sorted_topk_equals_oracle(setup, st):
    sh, qs, gix = setup
    k = 40
    res = GpuIndexSearcher(gix).search_sorted(qs, SortFieldCollector(k, st))
    wd, wv, wc, wt = oracle_sorted(sh, qs, k, st)
    assert np.array_equal(res.counts, wc) and np.array_equal(res.total_hits, wt) and not res.relation.any()
    for q in range(len(qs)):
        n = wc[q]
        assert np.array_equal(res.docs[q, :n], wd[q, :n]), (q, res.docs[q, :8], wd[q, :8])
        assert np.array_equal(res.sort_values[q, :n], wv[q, :n]), q


@pytest.mark.parametrize("st", [SORTS[0], SORTS[3], SORTS[5], SORTS[8]], ids=["int-asc", "float-desc", "long-desc-last", "docid-desc"])
def test_sorted_search_after_pages(setup, st):
    sh, qs, gix = setup
    k = 25
    s = GpuIndexSearcher(gix)
    p1 = s.search_sorted(qs, SortFieldCollector(k, st))
    after = [FieldDoc(int(p1.docs[q, k - 1]), int(p1.sort_values[q, k - 1])) if p1.counts[q] == k else None for q in range(len(qs))]
    sel = [q for q in range(len(qs)) if after[q] is not None]
    p2 = s.search_sorted([qs[q] for q in sel], SortFieldCollector(k, st), search_after=[after[q] for q in sel])
    wd, wv, wc, wt = oracle_sorted(sh, [qs[q] for q in sel], k, st, [after[q] for q in sel])
    assert np.array_equal(p2.counts, wc)
    for i in range(len(sel)):
        assert np.array_equal(p2.docs[i, :wc[i]], wd[i, :wc[i]]) and np.array_equal(p2.sort_values[i, :wc[i]], wv[i, :wc[i]])
    full = s.search_sorted([qs[q] for q in sel], SortFieldCollector(2 * k, st))
    for i, q in enumerate(sel):   # no overlap, no gap
        assert np.array_equal(np.concatenate([p1.docs[q, :k], p2.docs[i, :p2.counts[i]]]), full.docs[i, :k + p2.counts[i]])
