"""Host-side mirror of the reference interface: query flattening, QueryNodeMapper rules, synthetic inputs."""
import numpy as np
import pytest

from nrtsearch_b200 import index as ix
from nrtsearch_b200.search import (BooleanQuery, BoostQuery, MatchAllDocsQuery, Occur, RangeQuery, ScoreDoc, TermQuery,
                                   boolean_query_from_proto, compile_queries)
from nrtsearch_b200 import NrtGpuUnsupported


def test_boolean_query_from_proto_rules():
    # QueryNodeMapper.java:257-283
    q = boolean_query_from_proto([])
    assert len(q.clauses) == 1 and isinstance(q.clauses[0].query, MatchAllDocsQuery) and q.clauses[0].occur == Occur.MUST
    q = boolean_query_from_proto([(TermQuery(1), Occur.MUST_NOT), (TermQuery(2), Occur.MUST_NOT)])
    assert len(q.clauses) == 3 and q.clauses[-1].occur == Occur.FILTER
    q = boolean_query_from_proto([(TermQuery(1), Occur.MUST_NOT), (TermQuery(2), Occur.SHOULD)], 1)
    assert len(q.clauses) == 2 and q.minimum_number_should_match == 1


def test_compile_flattens_boosts_in_float_outermost_first():
    q = BoostQuery(BooleanQuery().add(BoostQuery(TermQuery(7), 1.1), Occur.SHOULD).add(RangeQuery(0, 5, 9), Occur.FILTER), 1.3)
    carr, ncl, qarr, nq = compile_queries([q, TermQuery(3)], [None, ScoreDoc(42, 1.5)])
    assert (ncl, nq) == (3, 2)
    assert np.float32(carr[0].boost) == np.float32(np.float32(1.3) * np.float32(1.1))
    assert (carr[0].occur, carr[0].kind, carr[0].id) == (0, 0, 7)
    assert (carr[1].occur, carr[1].kind, carr[1].lo, carr[1].hi) == (2, 1, 5, 9)
    assert (qarr[0].clause_begin, qarr[0].clause_end, qarr[0].has_after) == (0, 2, 0)
    assert (carr[2].occur, qarr[1].clause_begin, qarr[1].clause_end) == (1, 2, 3)   # bare leaf = single MUST
    assert (qarr[1].has_after, qarr[1].after_doc, qarr[1].after_score) == (1, 42, 1.5)


def test_compile_rejects_what_the_gpu_path_does_not_cover():
    with pytest.raises(NrtGpuUnsupported):
        compile_queries([BooleanQuery().add(BooleanQuery(), Occur.MUST)])
    with pytest.raises(ValueError, match="Boost must be a positive number"):
        compile_queries([BoostQuery(TermQuery(1), -1.0)])


def test_synth_corpus_is_deterministic_and_shardable(built):
    a = ix.synth_text_shard(6000, 500)
    b = ix.synth_text_shard(6000, 500)
    assert np.array_equal(a.post_docs, b.post_docs) and np.array_equal(a.post_freqs, b.post_freqs)
    assert a.fields[0].sum_total_term_freq == int(a.post_freqs.sum())
    # postings are doc-sorted per term, tf >= 1, norms = intToByte4(length)
    for t in (0, 1, 50, 499):
        seg = a.post_docs[a.term_off[t]:a.term_off[t + 1]]
        assert (np.diff(seg) > 0).all()
    lengths = np.bincount(a.post_docs, weights=a.post_freqs, minlength=6000).astype(int)
    import oracle
    assert all(a.fields[0].norms[d] == oracle.int_to_byte4(int(lengths[d])) for d in range(0, 6000, 97))
    # a doc-range shard generated on its own equals the slice of the whole corpus
    lo = ix.synth_text_shard(3000, 500, doc_begin=0)
    hi = ix.synth_text_shard(3000, 500, doc_begin=3000)
    sub = a.doc_range(3000, 6000)
    assert np.array_equal(hi.post_docs, sub.post_docs) and np.array_equal(hi.post_freqs, sub.post_freqs)
    assert hi.doc_base == 3000 and np.array_equal(np.diff(lo.term_off) + np.diff(hi.term_off), np.diff(a.term_off))
    assert np.array_equal(ix.synth_int_column(3000, doc_begin=3000), ix.synth_int_column(6000)[3000:])
    v = ix.synth_vectors(64, 8)
    assert np.array_equal(ix.synth_vectors(32, 8, row_begin=32), v[32:])
    assert abs(float(ix.synth_vectors(20000, 8).mean())) < 0.02


def test_query_terms_are_distinct_and_in_range(built):
    t = ix.synth_query_terms(1024, 3, 1_000_000)
    assert t.min() >= 10 and t.max() < 10_000
    assert all(len(set(r)) == 3 for r in t.tolist())


def test_knn_query_validation_messages():
    # VectorFieldDef.getKnnQuery :413-421
    import pytest
    from nrtsearch_b200.search import KnnQuery
    with pytest.raises(ValueError, match="Vector search k must be >= 1"):
        KnnQuery(0, 10).validate()
    with pytest.raises(ValueError, match="numCandidates must be >= k"):
        KnnQuery(10, 5).validate()
    with pytest.raises(ValueError, match="numCandidates > 10000"):
        KnnQuery(10, 10001).validate()
    KnnQuery(10, 100).validate()


def test_sort_missing_values_and_sortable_floats():
    # IntFieldDef.java:103, LongFieldDef.java:103, FloatFieldDef.java:105, DoubleFieldDef.java:105; NumericUtils sortable encodings
    from nrtsearch_b200.search import SortType, double_to_sortable_long, float_to_sortable_int
    assert SortType(0, False, True, "int").missing_value() == 2**31 - 1 and SortType(0, True, False, "int").missing_value() == -(2**31)
    assert SortType(0, False, True, "long").missing_value() == 2**63 - 1
    xs = [-float("inf"), -3.5, -0.0, 0.0, 1e-30, 2.0, float("inf")]
    assert [float_to_sortable_int(x) for x in xs] == sorted(float_to_sortable_int(x) for x in xs)
    assert [double_to_sortable_long(x) for x in xs] == sorted(double_to_sortable_long(x) for x in xs)
    assert SortType(0, False, True, "float").missing_value() == float_to_sortable_int(float("inf"))


def test_packed_record_layout_roundtrip():
    from nrtsearch_b200.shards import packed_words, unpack_record
    nq, k = 5, 3
    w = packed_words(nq, k)
    assert w == ((2 * nq * k + 2 * nq + 1) & ~1) + 2 * nq
    r = np.arange(w, dtype=np.int32)
    d, s, c, f, t = unpack_record(r, nq, k)
    assert d.shape == (nq, k) and s.shape == (nq, k) and len(c) == nq and len(f) == nq and len(t) == nq
    assert d[0, 0] == 0 and c[0] == 2 * nq * k and f[0] == 2 * nq * k + nq
