"""Deadline and terminateAfter behind the C ABI: SearchCutoffWrapper.java:164-202 and
TerminateAfterWrapper.java:85-162 semantics on the CUDA path.

What is pinned and how: the reference's slices share one AtomicInteger (TerminateAfterWrapper.java:150) and one wall
clock, so WHICH documents are collected before the cut is timing dependent in the reference as well. Parity is therefore
the wrappers' contract, checked against the oracle's sequential restatement:
  * terminatedEarly is set iff more than terminateAfter docs match (exact parity with the oracle);
  * totalHits is in [terminateAfter, min(matches, terminateAfterMaxRecallCount)] with relation GREATER_THAN_OR_EQUAL_TO;
  * every returned hit is a true match carrying its exact score (checked against the exhaustive oracle);
  * a query that did not terminate early returns exactly the unlimited result;
  * a passed deadline returns partial results + hitTimeout, or fails with CollectionTimeoutException."""
import numpy as np
import pytest

import oracle
from helpers import assert_same_hits
from nrtsearch_b200 import index as ix
from nrtsearch_b200._native import CollectionTimeoutException
from nrtsearch_b200.search import (BooleanQuery, GpuIndex, GpuIndexSearcher, Occur, RelevanceCollector, TermQuery,
                                   compile_queries)

pytestmark = pytest.mark.gpu
INT_MAX = 2**31 - 1


@pytest.fixture(scope="module")
def setup(gpu_ctx):
    sh = ix.synth_text_shard(1_300_000, 30_000, min_len=4, poisson_mean=14.0)   # 3 slices
    rng = np.random.default_rng(31)
    qs = []
    for i in range(96):
        ranks = np.unique(np.floor(10 ** rng.uniform(0.3, 4.2, size=1 + i % 3)).astype(np.int64).clip(1, 29_999))
        q = BooleanQuery()
        for r in ranks:
            q.add(TermQuery(int(r)), Occur.SHOULD if i % 4 else Occur.MUST)
        qs.append(q)
    gix = GpuIndex(gpu_ctx, sh)
    yield sh, qs, gix
    gix.close()


def test_terminate_after_contract(setup):
    sh, qs, gix = setup
    T, R, k = 2000, 5000, 50
    s = GpuIndexSearcher(gix)
    res = s.search_batch(qs, RelevanceCollector(k, INT_MAX, terminate_after=T, terminate_after_max_recall_count=R))
    full = s.search_batch(qs, RelevanceCollector(k, INT_MAX))
    carr, ncl, qarr, nq = compile_queries(qs)
    oix = oracle.OracleIndex(sh)
    od, os_, oc, ot, orel, oterm = oracle.search_terminate_after(oix, carr, ncl, qarr, nq, k, T, R)
    matches = oracle.search_compiled(oix, carr, ncl, qarr, nq, k)[3]
    assert np.array_equal(res.terminated_early, oterm), "terminatedEarly differs from the wrapper's rule (matches > terminateAfter)"
    assert np.array_equal(oterm != 0, matches > T)
    t = res.terminated_early != 0
    assert t.any() and (~t).any()
    assert (res.relation[t] == 1).all()
    assert (res.total_hits[t] >= T).all() and (res.total_hits[t] <= np.minimum(matches[t], R)).all()
    assert (ot[t] == np.minimum(matches[t], R)).all()   # the sequential oracle counts up to the recall cap
    # unaffected queries: identical to the unlimited search (and to the oracle under the wrapper)
    for q in np.nonzero(~t)[0]:
        n = res.counts[q]
        assert n == full.counts[q] == oc[q]
        assert np.array_equal(res.docs[q, :n], full.docs[q, :n]) and np.array_equal(res.docs[q, :n], od[q, :n])
        assert np.array_equal(res.scores[q, :n].view(np.uint32), os_[q, :n].view(np.uint32))
        assert res.total_hits[q] == matches[q]
    # terminated queries: every hit is a true match with its exact score (scores of all matches from a deep oracle run)
    sel = [int(q) for q in np.nonzero(t)[0][:12]]
    carr2, ncl2, qarr2, nq2 = compile_queries([qs[q] for q in sel])
    deep = 4096
    dd, ds, dc, dt, _ = oracle.search_compiled(oix, carr2, ncl2, qarr2, nq2, deep)
    for i, q in enumerate(sel):
        truth = {int(d): s_ for d, s_ in zip(dd[i, :dc[i]], ds[i, :dc[i]])}
        worst = ds[i, dc[i] - 1]
        for d, sc in zip(res.docs[q, :res.counts[q]], res.scores[q, :res.counts[q]]):
            if int(d) in truth:
                assert np.float32(sc).view(np.uint32) == np.float32(truth[int(d)]).view(np.uint32)
            else:
                assert dc[i] == deep and sc <= worst   # outside the deep list: must rank below it


def test_deadline_partial_results_and_exception(setup):
    sh, qs, gix = setup
    s = GpuIndexSearcher(gix)
    k = 20
    full = s.search_batch(qs, RelevanceCollector(k, INT_MAX))
    # a generous deadline changes nothing
    ok = s.search_batch(qs, RelevanceCollector(k, INT_MAX, timeout_sec=120.0))
    assert not ok.hit_timeout.any()
    assert np.array_equal(ok.docs, full.docs) and np.array_equal(ok.total_hits, full.total_hits) and not ok.relation.any()
    # the request had already used up its budget before the call: every work item is skipped
    late = s.search_batch(qs, RelevanceCollector(k, INT_MAX, timeout_sec=0.5, elapsed_sec=1.0))
    assert late.hit_timeout.all() and (late.relation == 1).all() and (late.counts == 0).all() and (late.total_hits == 0).all()
    # a deadline that falls inside the batch (device clock, checked at work-item boundaries): whatever was collected
    # before it is returned, flagged, and is a prefix-consistent subset (never more hits than the full search)
    mid = s.search_batch(qs * 8, RelevanceCollector(k, INT_MAX, timeout_sec=20e-6))
    to = mid.hit_timeout != 0
    assert (mid.relation[to] == 1).all() and (mid.total_hits <= np.tile(full.total_hits, 8)).all()
    assert np.array_equal(mid.docs[~to], np.tile(full.docs, (8, 1))[~to])
    with pytest.raises(CollectionTimeoutException, match="Search collection exceeded timeout of"):
        s.search_batch(qs, RelevanceCollector(k, INT_MAX, timeout_sec=0.5, elapsed_sec=1.0, disallow_partial_results=True))
