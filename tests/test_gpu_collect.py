"""Either side of the top-k on the CUDA path: additional collectors (terms / min / max / sum over every matching doc),
the second pass of QueryRescorer on the device, and the fetch phase on doc-value columns. Oracle: the exhaustive CPU
evaluation's match stream + numpy (the arithmetic of the reference's collectors is a count / compare / add per doc)."""
import numpy as np
import pytest

import oracle
from nrtsearch_b200 import index as ix
from nrtsearch_b200.search import (BooleanQuery, GpuIndex, GpuIndexSearcher, MatchAllDocsQuery, MaxCollector, MinCollector, Occur,
                                   RangeQuery, RelevanceCollector, SumCollector, TermQuery, TermsCollector, compile_queries,
                                   float_to_sortable_int)

pytestmark = pytest.mark.gpu
INT_MAX = 2**31 - 1


@pytest.fixture(scope="module")
def setup(gpu_ctx):
    n = 200_000
    sh = ix.synth_text_shard(n, 6_000, min_len=6, poisson_mean=24.0)
    rng = np.random.default_rng(51)
    category = rng.integers(0, 300, n).astype(np.int64)                    # categorical int field
    price = ix.synth_int_column(n, value_range=100_000)
    rating_f = rng.normal(3.0, 1.0, n).astype(np.float32)
    rating = np.array([float_to_sortable_int(x) for x in np.unique(rating_f)], np.int64)[np.searchsorted(np.unique(rating_f), rating_f)]
    has_rating = (rng.random(n) < 0.8).astype(np.uint8)
    sh.columns = [category, price, rating]
    sh.column_has = [None, None, has_rating]
    terms = ix.synth_query_terms(24, 3, 6_000, log10_lo=0.3, log10_hi=3.0)
    qs = []
    for i, t in enumerate(terms):
        if i % 4 == 0:
            qs.append(BooleanQuery().add(TermQuery(int(t[0])), Occur.MUST).add(RangeQuery(1, 1_000, 60_000), Occur.FILTER))
        elif i % 4 == 1:
            qs.append(MatchAllDocsQuery() if i % 8 == 1 else BooleanQuery().add(TermQuery(int(t[0])), Occur.SHOULD).add(TermQuery(int(t[1])), Occur.MUST_NOT))
        else:
            q = BooleanQuery()
            for x in t:
                q.add(TermQuery(int(x)), Occur.SHOULD)
            qs.append(q)
    gix = GpuIndex(gpu_ctx, sh)
    yield sh, qs, gix, rating_f
    gix.close()


def test_additional_collectors(setup):
    sh, qs, gix, rating_f = setup
    s = GpuIndexSearcher(gix)
    adds = [TermsCollector(0, 10), TermsCollector(0, 7, order_desc=False), MinCollector(1, "int"), MaxCollector(1, "int"), SumCollector(1, "int"),
            MaxCollector(2, "float"), SumCollector(2, "float"), TermsCollector(1, 5)]
    res, outs = s.search_with_collectors(qs, RelevanceCollector(20, 1000), adds)
    plain = s.search_batch(qs, RelevanceCollector(20, INT_MAX))
    assert np.array_equal(res.docs, plain.docs) and np.array_equal(res.total_hits, plain.total_hits)   # hits unchanged, counts exact
    carr, ncl, qarr, nq = compile_queries(qs)
    oix = oracle.OracleIndex(sh)
    cat, price, has_r = sh.columns[0], sh.columns[1], sh.column_has[2]
    for q in range(nq):
        m = oracle.match_bitmap(oix, carr, qarr, q).astype(bool)
        assert m.sum() == plain.total_hits[q]
        for oi, col in ((0, cat), (1, cat), (7, price)):
            vals, cnts = np.unique(col[m], return_counts=True)
            o = outs[oi]
            size, desc = adds[oi].size, adds[oi].order_desc
            n = o["n"][q]
            assert n == min(size, len(vals)) and o["total_buckets"][q] == len(vals)
            got = dict(zip(o["keys"][q, :n].tolist(), o["counts"][q, :n].tolist()))
            truth = dict(zip(vals.tolist(), cnts.tolist()))
            assert all(truth.get(k) == c for k, c in got.items()), "bucket count differs"
            order = sorted(cnts.tolist(), reverse=desc)[:n]
            assert o["counts"][q, :n].tolist() == order                       # the right multiset of counts, in order
            assert o["other_counts"][q] == int(cnts.sum()) - sum(order)
        if m.any():
            assert outs[2][q] == float(price[m].min()) and outs[3][q] == float(price[m].max()) and outs[4][q] == float(price[m].sum())
        else:
            assert outs[2][q] == np.finfo(np.float64).max and outs[3][q] == -np.finfo(np.float64).max and outs[4][q] == 0.0
        mr = m & (has_r != 0)
        if mr.any():
            assert outs[5][q] == float(rating_f[mr].max())
            np.testing.assert_allclose(outs[6][q], rating_f[mr].astype(np.float64).sum(), rtol=1e-12)


def test_query_rescorer_second_pass_and_fetch(setup):
    sh, qs, gix, _ = setup
    s = GpuIndexSearcher(gix)
    first = s.search_batch(qs, RelevanceCollector(60, INT_MAX))
    # second query: a different boolean per request, evaluated on the first-pass hits only
    terms = ix.synth_query_terms(len(qs), 2, 6_000, seed=99, log10_lo=0.3, log10_hi=2.5)
    second = [BooleanQuery().add(TermQuery(int(t[0])), Occur.SHOULD).add(TermQuery(int(t[1])), Occur.SHOULD)
              .add(RangeQuery(1, 0, 80_000), Occur.FILTER) for t in terms]
    m, sc = s.score_docs(second, first.docs, first.counts)
    carr, ncl, qarr, nq = compile_queries(second)
    oix = oracle.OracleIndex(sh)
    wm, ws = oracle.score_docs(oix, carr, qarr, nq, first.docs, first.counts)
    assert np.array_equal(m, wm) and np.array_equal(sc.view(np.uint32), ws.view(np.uint32)) and wm.any()
    # the whole QueryRescore on the device vs oracle second pass + oracle combine (window 40 of 60 hits)
    d, r, c = s.rescore_query(second, first.docs, first.scores, first.counts, 40, 1.0, 2.5)
    for q in range(nq):
        n = first.counts[q]
        od, os_ = oracle.rescore_combine(first.docs[q, :n], first.scores[q, :n], wm[q, :n], ws[q, :n], 1.0, 2.5)
        keep = min(n, 40)
        assert c[q] == keep and np.array_equal(d[q, :keep], od[:keep]) and np.array_equal(r[q, :keep].view(np.uint32), os_[:keep].view(np.uint32))
    # fetch phase: doc values of the final hits
    flat = first.docs[first.counts > 0][:, 0]
    vals, has = s.fetch_columns([1, 2, 0], flat)
    assert np.array_equal(vals[0], sh.columns[1][flat]) and np.array_equal(vals[2], sh.columns[0][flat]) and has[0].all()
    assert np.array_equal(has[1], sh.column_has[2][flat]) and np.array_equal(vals[1][has[1] != 0], sh.columns[2][flat][has[1] != 0])


def test_micro_batcher_matches_direct_search(setup):
    """96 handler threads submit one query each (the reference's one-query-per-RPC API); the native batcher groups them.
    Every caller gets exactly what a direct search returns, bad requests fail alone, and requests really share batches."""
    import threading
    from nrtsearch_b200 import NrtGpuError
    from nrtsearch_b200.search import GpuBatcher
    sh, qs, gix, _ = setup
    direct = GpuIndexSearcher(gix).search_batch(qs, RelevanceCollector(15, INT_MAX))
    b = GpuBatcher(gix, max_batch=64, max_wait_us=20_000)
    results, errors = {}, {}

    def worker(i):
        try:
            if i % 17 == 5:
                results[i] = b.submit(TermQuery(10**8), RelevanceCollector(15, INT_MAX))   # out-of-range term: must fail alone
            else:
                results[i] = b.submit(qs[i % len(qs)], RelevanceCollector(15, INT_MAX))
        except NrtGpuError as e:
            errors[i] = e
    ts = [threading.Thread(target=worker, args=(i,)) for i in range(96)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    st = b.stats()
    b.close()
    bad = {i for i in range(96) if i % 17 == 5}
    assert set(errors) == bad and all("term id out of range" in str(e) for e in errors.values())
    for i in set(range(96)) - bad:
        td, diag = results[i]
        q = i % len(qs)
        n = direct.counts[q]
        assert [sd.doc for sd in td.score_docs] == direct.docs[q, :n].tolist()
        assert np.array_equal(np.array([sd.score for sd in td.score_docs], np.float32).view(np.uint32), direct.scores[q, :n].view(np.uint32))
        assert td.total_hits.value == direct.total_hits[q] and diag.batch_size >= 1 and diag.search_ms > 0
    assert st["requests"] >= 96 - len(bad) and st["batches"] < st["requests"], st   # requests shared batches
