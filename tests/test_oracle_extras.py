"""The oracle's functions around the top-k -- sort-by-field, terminateAfter, second-pass scoring, score-order blending -- are
the checkers of the matching CUDA paths (tests/test_gpu_sort.py, test_gpu_limits.py, test_gpu_collect.py,
test_gpu_hybrid.py). Here each is pinned on the CPU against an independent numpy restatement of the reference semantics
(TopFieldCollector: SortFieldCollector.java:44-105; TerminateAfterWrapper.java:150-162; QueryRescorer's second pass:
QueryRescore.java:39-57; WeightedScoreOrderBlenderOperation), so that the GPU tests compare against something that was
itself compared."""
import numpy as np

import oracle
from nrtsearch_b200 import index as ix
from nrtsearch_b200.search import BooleanQuery, Occur, RangeQuery, ScoreDoc, TermQuery, compile_queries

INT_MAX = 2**31 - 1


def corpus(n=40_000, vocab=2_000, seed=9):
    sh = ix.synth_text_shard(n, vocab, min_len=4, poisson_mean=12.0)
    rng = np.random.default_rng(seed)
    col = rng.integers(-50, 50, n).astype(np.int64)           # many ties
    has = (rng.random(n) > 0.2).astype(np.uint8)               # 20 % of the docs have no value
    sh.columns = [col]
    sh.column_has = [has]
    return sh


def some_queries(vocab, n=24, seed=2):
    terms = ix.synth_query_terms(n, 2, vocab, seed=seed, log10_lo=0.3, log10_hi=3.0)
    qs = []
    for i, t in enumerate(terms):
        q = BooleanQuery()
        q.add(TermQuery(int(t[0])), Occur.SHOULD if i % 2 else Occur.MUST)
        q.add(TermQuery(int(t[1])), Occur.SHOULD)
        if i % 3 == 0:
            q.add(RangeQuery(0, -20, 30), Occur.FILTER)
        qs.append(q)
    return qs


def test_sorted_search_matches_numpy_order():
    sh = corpus()
    oix = oracle.OracleIndex(sh)
    qs = some_queries(2_000)
    carr, ncl, qarr, nq = compile_queries(qs)
    col, has = sh.columns[0], sh.column_has[0].astype(bool)
    for reverse in (False, True):
        for missing in (-(2**63), 2**63 - 1, 7):
            d, v, c, t = oracle.search_sorted(oix, carr, ncl, qarr, nq, 25, 1, 0, reverse, missing)
            for q in range(nq):
                m = np.nonzero(oracle.match_bitmap(oix, carr, qarr, q))[0]
                vals = np.where(has[m], col[m], missing).astype(object)        # object: Long.MIN/MAX stay exact
                order = sorted(range(len(m)), key=lambda i: ((-vals[i]) if reverse else vals[i], m[i]))[:25]
                assert t[q] == len(m)
                assert list(d[q, :c[q]]) == [int(m[i]) for i in order], (q, reverse, missing)
                assert list(v[q, :c[q]]) == [int(vals[i]) for i in order]
    # searchAfter on (value, doc): the page after the 10th hit starts at the 11th
    d, v, c, t = oracle.search_sorted(oix, carr, ncl, qarr, nq, 25, 1, 0, False, 7)
    sel = [q for q in range(nq) if c[q] == 25]
    after = [ScoreDoc(int(d[q, 9]), 0.0) for q in sel]
    carr2, ncl2, qarr2, nq2 = compile_queries([qs[q] for q in sel], after)
    d2, v2, c2, _ = oracle.search_sorted(oix, carr2, ncl2, qarr2, nq2, 15, 1, 0, False, 7, after_values=[int(v[q, 9]) for q in sel])
    for i, q in enumerate(sel):
        assert list(d2[i, :15]) == list(d[q, 10:25])


def test_terminate_after_is_sequential_in_doc_order():
    sh = corpus()
    oix = oracle.OracleIndex(sh)
    qs = some_queries(2_000, n=16, seed=4)
    carr, ncl, qarr, nq = compile_queries(qs)
    full = oracle.search_compiled(oix, carr, ncl, qarr, nq, 1000)      # every match with its score (few thousand docs at most)
    for ta in (5, 60):
        d, s, c, t, rel, term = oracle.search_terminate_after(oix, carr, ncl, qarr, nq, 10, ta)
        for q in range(nq):
            m = np.nonzero(oracle.match_bitmap(oix, carr, qarr, q))[0]
            if len(m) <= ta:                                            # never reached: the plain search
                assert not term[q] and t[q] == len(m)
                assert list(d[q, :c[q]]) == list(full[0][q, :min(10, full[2][q])])
                continue
            assert term[q] and rel[q] == 1 and t[q] == ta              # TerminateAfterWrapper.java:85-90, 150-158
            first = set(int(x) for x in m[:ta])                         # the first `ta` matches in doc order were collected
            if full[2][q] < len(m):
                continue                                               # (more matches than the reference page holds: skip the order check)
            score_of = {int(full[0][q, i]): full[1][q, i] for i in range(full[2][q])}
            want = sorted(first, key=lambda x: (-float(score_of[x]), x))[:10]
            assert list(d[q, :c[q]]) == want


def test_score_docs_reproduces_the_search_scores():
    sh = corpus()
    oix = oracle.OracleIndex(sh)
    qs = some_queries(2_000, n=20, seed=6)
    carr, ncl, qarr, nq = compile_queries(qs)
    d, s, c, t, r = oracle.search_compiled(oix, carr, ncl, qarr, nq, 30)
    m, s2 = oracle.score_docs(oix, carr, qarr, nq, d, c)
    for q in range(nq):
        assert m[q, :c[q]].all() and not m[q, c[q]:].any()
        assert np.array_equal(s2[q, :c[q]].view(np.uint32), s[q, :c[q]].view(np.uint32))
    # a doc the query does not match: no match, score 0
    probe = np.zeros((nq, 1), np.int32)
    for q in range(nq):
        probe[q, 0] = int(np.nonzero(oracle.match_bitmap(oix, carr, qarr, q) == 0)[0][0])
    m, s2 = oracle.score_docs(oix, carr, qarr, nq, probe)
    assert not m.any() and not s2.any()


def test_score_order_blenders_against_a_dict():
    rng = np.random.default_rng(11)
    for mode, fold in ((1, max), (2, sum), (3, None)):
        for _ in range(20):
            R, top_in = 3, 12
            docs = np.stack([rng.choice(40, top_in, replace=False) for _ in range(R)]).astype(np.int32)
            scores = np.sort(rng.random((R, top_in)).astype(np.float32), axis=1)[:, ::-1]
            counts = rng.integers(0, top_in + 1, R).astype(np.int32)
            boosts = np.array([1.0, 0.5, 2.0], np.float32)
            per_doc = {}
            for r in range(R):
                for i in range(counts[r]):
                    per_doc.setdefault(int(docs[r, i]), []).append(np.float32(scores[r, i] * boosts[r]))
            want = {}
            for doc, xs in per_doc.items():
                if mode == 1:
                    want[doc] = max(xs)
                else:
                    acc = np.float32(0.0)
                    for x in xs:
                        acc = np.float32(acc + x)                      # float accumulation in retriever order
                    want[doc] = acc if mode == 2 else np.float32(acc / np.float32(len(xs)))
            order = sorted(want, key=lambda x: (-float(want[x]), x))[:10]
            od, os_, total = oracle.blend_scores(mode, docs, scores, counts, boosts, 10)
            assert total == len(want)
            assert list(od) == order
            np.testing.assert_allclose(os_, [want[x] for x in order], rtol=1e-6)
