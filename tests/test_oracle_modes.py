"""Oracle self-consistency: pruned (MAXSCORE) == exhaustive; double clause sums are order independent on
the synthetic corpora (the assumption that makes bit-exact parity on a parallel machine possible);
doc-range shards + TopDocs.merge reproduce the single-index result."""
import numpy as np

import oracle
from nrtsearch_b200 import index as ix
from nrtsearch_b200.search import BooleanQuery, Occur, TermQuery, compile_queries

INT_MAX = 2**31 - 1


def disj(terms):
    q = BooleanQuery()
    for t in terms:
        q.add(TermQuery(int(t)), Occur.SHOULD)
    return q


def test_pruned_equals_exhaustive(built):
    sh = ix.synth_text_shard(120_000, 30_000)
    terms = ix.synth_query_terms(200, 3, 30_000, log10_lo=0.3, log10_hi=3.8)
    carr, ncl, qarr, nq = compile_queries([disj(t) for t in terms])
    oix = oracle.OracleIndex(sh, with_impacts=True)
    ex = oracle.search_compiled(oix, carr, ncl, qarr, nq, 100)
    pr = oracle.search_compiled(oix, carr, ncl, qarr, nq, 100, total_hits_threshold=1000, mode=1)
    assert np.array_equal(ex[0], pr[0]) and np.array_equal(ex[1].view(np.uint32), pr[1].view(np.uint32))
    assert np.array_equal(ex[2], pr[2])
    assert (pr[3] <= ex[3]).all() and (pr[3][pr[4] == 0] == ex[3][pr[4] == 0]).all()
    assert (pr[4] == 1).any()     # pruning did happen
    assert (ex[4] == 0).all()


def test_double_sum_is_order_independent(built):
    sh = ix.synth_text_shard(60_000, 10_000)
    terms = ix.synth_query_terms(64, 3, 10_000, log10_lo=0.3, log10_hi=3.0)
    oix = oracle.OracleIndex(sh)
    a = oracle.search_compiled(oix, *compile_queries([disj(t) for t in terms]), 100)
    b = oracle.search_compiled(oix, *compile_queries([disj(t[::-1]) for t in terms]), 100)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_doc_range_shards_merge_to_the_whole(built):
    sh = ix.synth_text_shard(50_000, 5_000)
    sh.term_df = np.diff(sh.term_off).astype(np.int64)
    terms = ix.synth_query_terms(32, 3, 5_000, log10_lo=0.3, log10_hi=3.0)
    cq = compile_queries([disj(t) for t in terms])
    whole = oracle.search_compiled(oracle.OracleIndex(sh), *cq, 20)
    parts = [sh.doc_range(0, 20_000), sh.doc_range(20_000, 37_000), sh.doc_range(37_000, 50_000)]
    res = [oracle.search_compiled(oracle.OracleIndex(p), *cq, 20) for p in parts]
    d, s, c = oracle.merge_topk(np.stack([r[0] for r in res]), np.stack([r[1] for r in res]), np.stack([r[2] for r in res]), 20)
    assert np.array_equal(d, whole[0]) and np.array_equal(s.view(np.uint32), whole[1].view(np.uint32))
    assert np.array_equal(sum(r[3] for r in res), whole[3])
