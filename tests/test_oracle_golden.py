"""Pins the CPU oracle to every known answer the reference's own tests hold for the hot path
(SURVEY.md 8c). The constants below are copied from the reference TEST EXPECTATIONS (not code):
  MultiFunctionScoreQueryTest.java:139, SearchStateTest.java:117, QueryTest.java:1003-1018,
  SimilarityTest.java:114-120, docker-compose-config/{docs.csv,search.json}, QueryTest.java:932-975,
  VectorFieldDefTest.java:2433-2578, MultiRetrieverSearchTest.java:410-437."""
import numpy as np
import pytest

import oracle
from helpers import shard_from_token_docs

INT_MAX = 2**31 - 1


def f32(x):
    return float(np.float32(x))


def term_search(sh, vocab, field, tokens, occur=0, top_k=10, **kw):
    clauses = [(occur, 0, vocab[(field, t)], 1.0, 0, 0) for t in tokens]
    q = [(0, len(clauses), 0, 0, 0, 0.0)]
    return oracle.search(oracle.OracleIndex(sh), clauses, q, top_k, **kw)


def test_smallfloat_roundtrip():
    # lengths <= 40 are exact; 41 -> 40, 100 -> 96, 1000 -> 984 (4-bit mantissa)
    for i in range(41):
        assert oracle.byte4_to_int(oracle.int_to_byte4(i)) == i
    assert oracle.byte4_to_int(oracle.int_to_byte4(41)) == 40
    assert oracle.byte4_to_int(oracle.int_to_byte4(100)) == 96
    assert oracle.byte4_to_int(oracle.int_to_byte4(1000)) == 984
    assert oracle.int_to_byte4(2**31 - 1) == 255
    prev = -1
    for b in range(256):  # decode is monotone
        v = oracle.byte4_to_int(b)
        assert v > prev
        prev = v


def test_bm25_multifunction_corpus_bitexact():
    # MultiFunctionScoreQueryTest corpus (:60-113): MatchQuery text_field:"Document2" -> docs 2, 4
    docs = ["Document1 with none of filter terms", "Document2 with term1 filter term",
            "Document1 with term2 filter term", "Document2 with both term1 and term2 filter terms"]
    sh, vocab = shard_from_token_docs([[d.lower().split() for d in docs]])
    d, s, c, t, r = term_search(sh, vocab, 0, ["document2"])
    assert c[0] == 2 and list(d[0, :2]) == [1, 3] and t[0] == 2
    assert float(s[0, 0]) == 0.33812057971954346
    assert float(s[0, 1]) == 0.27725890278816223


def test_bm25_search_state_corpus():
    # SearchStateTest (:43-66, :117): vendor_name:vendor over {"first vendor", "second vendor review"}
    sh, vocab = shard_from_token_docs([["first vendor".split(), "second vendor review".split()]])
    d, s, c, t, r = term_search(sh, vocab, 0, ["vendor"])
    assert list(d[0, :2]) == [0, 1]
    assert abs(float(s[0, 1]) - 0.0766057) < 1e-7   # lastScore of lastDocId 1


def test_bm25_explain_constants():
    # QueryTest.java:1003-1018: idf(n=1,N=2)=0.6931472, idf(n=2,N=2)=0.18232156, tf=0.45454544 at dl=avgdl=4
    idf1, idf2 = oracle.bm25_idf(1, 2), oracle.bm25_idf(2, 2)
    assert f32(idf1) == f32(0.6931472) and f32(idf2) == f32(0.18232156)
    idf_sum = np.float32(np.float64(idf1) + np.float64(idf2))   # phrase: (float) sum of idfs
    assert f32(idf_sum) == f32(0.87546873)
    cache = oracle.bm25_cache(1.2, 0.75, 4.0)
    L = oracle.lib()
    import ctypes as C
    score = L.orc_bm25_score(C.c_float(idf_sum), C.c_float(1.0), oracle.int_to_byte4(4),
                             cache.ctypes.data_as(C.POINTER(C.c_float)))
    assert f32(score) == f32(0.3979403)
    tf = L.orc_bm25_score(C.c_float(1.0), C.c_float(1.0), oracle.int_to_byte4(4), cache.ctypes.data_as(C.POINTER(C.c_float)))
    assert abs(tf - 0.45454544) < 1e-7


def test_config1_docker_compose():
    # docker-compose-config/docs.csv + search.json: "vendor_name:first vendor" -> first SHOULD vendor SHOULD
    sh, vocab = shard_from_token_docs([["first vendor".split(), "second vendor".split()]])
    d, s, c, t, r = term_search(sh, vocab, 0, ["first", "vendor"])
    assert list(d[0, :2]) == [0, 1] and t[0] == 2
    assert f32(s[0, 0]) == f32(0.3979403)        # 0.31506687 + 0.08287343 summed in double
    assert abs(float(s[0, 1]) - 0.0828734) < 1e-7    # same term score SimilarityTest pins as 0.0828734


def test_similarity_test_components():
    # SimilarityTest.java:114-120: 12.12609 = bm25(first, tf=2)=0.43321696 + bm25(vendor)=0.0828734 + 0.5 + 11.11
    a = ["first", "vendor", "first", "again"]
    b = ["second", "vendor", "second", "again"]
    sh, vocab = shard_from_token_docs([[a, b]])
    d, s, c, t, r = term_search(sh, vocab, 0, ["first"])
    assert abs(float(s[0, 0]) - 0.43321696) < 1e-7
    d, s, c, t, r = term_search(sh, vocab, 0, ["vendor"])
    assert abs(float(s[0, 0]) - 0.0828734) < 1e-7
    d, s, c, t, r = term_search(sh, vocab, 0, ["first", "vendor"])
    assert abs(float(s[0, 0]) + 0.5 + 11.11 - 12.12609) < 1e-4
    assert abs(float(s[0, 1]) + 0.5 + 11.11 - 11.692873) < 1e-4


def test_boost_linearity_power_of_two():
    # QueryTest.java:1136-1160: boost*score == boosted (exact for boost = 2)
    sh, vocab = shard_from_token_docs([["first vendor".split(), "second vendor review".split()]])
    oix = oracle.OracleIndex(sh)
    t = vocab[(0, "vendor")]
    _, s1, *_ = oracle.search(oix, [(1, 0, t, 1.0, 0, 0)], [(0, 1, 0, 0, 0, 0.0)], 10)
    _, s2, *_ = oracle.search(oix, [(1, 0, t, 2.0, 0, 0)], [(0, 1, 0, 0, 0, 0.0)], 10)
    assert np.array_equal(s1 * np.float32(2.0), s2)


def test_range_query_hit_set():
    # QueryTest.testSearchRangeQuery (:932-975) over src/test/resources/addDocs.csv: each range matches doc "2" only
    count = np.array([3, 7], np.int64)
    long_field = np.array([12, 16], np.int64)
    sh, vocab = shard_from_token_docs([["first vendor".split(), "second vendor".split()]], columns=[count, long_field])
    oix = oracle.OracleIndex(sh)
    for col, lo, hi in ((0, 5, 10), (1, 15, 19)):
        d, s, c, t, r = oracle.search(oix, [(1, 1, col, 1.0, lo, hi)], [(0, 1, 0, 0, 0, 0.0)], 10)
        assert c[0] == 1 and d[0, 0] == 1 and t[0] == 1
        assert float(s[0, 0]) == 1.0   # constant-score query


def test_multivalued_range_matches_any_value():
    # IntFieldDefTest's "multi_stored" docs (:131-141): doc 1 holds {Integer.MIN_VALUE, 15}, doc 2 holds {1, 15}; a third doc
    # has no value. SortedNumericDocValuesRangeQuery: a doc matches when ANY value is inside the inclusive range.
    imin = -(2**31)
    vals = np.array([imin, 15, 1, 15], np.int64)
    off = np.array([0, 2, 4, 4], np.int64)
    sh, vocab = shard_from_token_docs([["a".split(), "a".split(), "a".split()]], columns=[vals])
    sh.column_offsets = [off]
    oix = oracle.OracleIndex(sh)
    for lo, hi, docs in ((15, 15, [0, 1]), (0, 10, [1]), (imin, imin, [0]), (2, 14, []), (16, 2**31 - 1, []), (imin, 2**31 - 1, [0, 1])):
        d, s, c, t, r = oracle.search(oix, [(1, 1, 0, 1.0, lo, hi)], [(0, 1, 0, 0, 0, 0.0)], 10)
        assert list(d[0, :c[0]]) == docs and t[0] == len(docs), (lo, hi)
    sub = sh.doc_range(1, 3)     # a doc-range shard re-bases the offsets
    assert list(sub.column_offsets[0]) == [0, 2, 2] and list(sub.columns[0]) == [1, 15]


def test_match_all_scores_one():
    # MultiFunctionScoreQueryTest.testNoFunctionsMatchAll (:117-126): every doc, score 1.0
    sh, vocab = shard_from_token_docs([["a b".split(), "c".split(), "d".split(), "e".split()]])
    d, s, c, t, r = oracle.search(oracle.OracleIndex(sh), [(1, 2, 0, 1.0, 0, 0)], [(0, 1, 0, 0, 0, 0.0)], 10)
    assert list(d[0, :4]) == [0, 1, 2, 3] and all(float(x) == 1.0 for x in s[0, :4])


def test_topk_order_ties_and_search_after():
    # LazyQueueTopScoreDocCollector.java:112,129-143: score desc, doc asc; searchAfter skips score>after or (== and doc<=after)
    docs = [["x", "pad"]] * 6   # six identical docs -> identical scores
    sh, vocab = shard_from_token_docs([docs])
    oix = oracle.OracleIndex(sh)
    t = vocab[(0, "x")]
    d, s, c, tot, r = oracle.search(oix, [(1, 0, t, 1.0, 0, 0)], [(0, 1, 0, 0, 0, 0.0)], 4)
    assert list(d[0, :4]) == [0, 1, 2, 3] and tot[0] == 6
    d2, s2, c2, tot2, _ = oracle.search(oix, [(1, 0, t, 1.0, 0, 0)], [(0, 1, 0, 1, 3, float(s[0, 3]))], 4)
    assert c2[0] == 2 and list(d2[0, :2]) == [4, 5] and tot2[0] == 6   # paging: no overlap, totalHits unchanged


def test_vector_score_formulas():
    # VectorFieldDefTest.java:2433-2578 / VectorFieldDef.java:664-673
    a = np.array([1.0, 2.0, 2.0], np.float32)
    b = np.array([0.5, -1.0, 2.0], np.float32)
    dot = float(np.dot(a.astype(np.float64), b.astype(np.float64)))
    d2 = float(((a.astype(np.float64) - b) ** 2).sum())
    cos = dot / np.sqrt(float((a.astype(np.float64) ** 2).sum()) * float((b.astype(np.float64) ** 2).sum()))
    assert abs(oracle.vector_score(a, b, 0) - 1.0 / (1.0 + d2)) < 1e-7
    assert abs(oracle.vector_score(a, b, 1) - (1.0 + dot) / 2.0) < 1e-6
    assert abs(oracle.vector_score(a, b, 2) - (1.0 + cos) / 2.0) < 1e-7
    assert abs(oracle.vector_score(a, b, 3) - (dot + 1.0)) < 1e-6
    assert abs(oracle.vector_score(a, -b, 3) - 1.0 / (1.0 + dot)) < 1e-7   # negative inner product branch


def test_rrf_blend_known_answers():
    # MultiRetrieverSearchTest.java:410-437: score = sum 1/(60+rank), 1e-5
    text = np.array([[7, 3, 5]], np.int32)[0]
    knn = np.array([[3, 9, 7]], np.int32)[0]
    docs = np.stack([text, knn])
    d, s, total = oracle.blend_rrf(docs, [3, 3], [1.0, 1.0], 60, 10)
    want = {7: 1 / 61 + 1 / 63, 3: 1 / 62 + 1 / 61, 5: 1 / 63, 9: 1 / 62}
    assert total == 4 and set(d.tolist()) == set(want)
    for doc, sc in zip(d, s):
        assert abs(float(sc) - want[int(doc)]) < 1e-5
    assert list(d[:2]) == [3, 7]
    # weights: boost/(k+rank)
    d, s, total = oracle.blend_rrf(docs, [3, 3], [2.0, 0.5], 60, 10)
    assert abs(float(s[list(d).index(7)]) - (2.0 / 61 + 0.5 / 63)) < 1e-6


def test_rescore_combine():
    # QueryRescore.combine (:39-46): qw*first + rw*second in double -> float; unmatched: qw*first
    docs = np.array([10, 11, 12], np.int32)
    scores = np.array([3.0, 2.0, 1.0], np.float32)
    d, s = oracle.rescore_combine(docs, scores, [1, 0, 1], [1.0, 0.0, 4.0], 1.0, 2.0)
    assert list(d) == [12, 10, 11] and [float(x) for x in s] == [9.0, 5.0, 2.0]
