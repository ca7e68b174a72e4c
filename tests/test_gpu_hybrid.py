"""Hybrid stages on the device vs the oracle: weighted RRF blend, QueryRescore.combine, cross-shard TopDocs.merge."""
import ctypes

import numpy as np
import pytest

import oracle
from nrtsearch_b200 import _native
from nrtsearch_b200.search import blend_rrf, rescore_combine

pytestmark = pytest.mark.gpu


def test_rrf_known_answers_and_oracle(gpu_ctx):
    # MultiRetrieverSearchTest.java:410-437: 1/(60+rank) sums
    docs = np.array([[[7, 3, 5]], [[3, 9, 7]]], np.int32)         # [R=2, nq=1, top_in=3]
    d, s, c, t = blend_rrf(gpu_ctx, docs, np.array([[3], [3]]), [1.0, 1.0], 60, 10)
    want = {7: 1 / 61 + 1 / 63, 3: 1 / 62 + 1 / 61, 5: 1 / 63, 9: 1 / 62}
    assert c[0] == 4 and t[0] == 4
    for doc, sc in zip(d[0, :4], s[0, :4]):
        assert abs(float(sc) - want[int(doc)]) < 1e-5
    rng = np.random.default_rng(1)
    R, nq, top_in = 3, 40, 100
    docs = np.stack([np.stack([rng.choice(500, top_in, replace=False) for _ in range(nq)]) for _ in range(R)]).astype(np.int32)
    counts = rng.integers(0, top_in + 1, (R, nq)).astype(np.int32)
    boosts = [1.0, 0.7, 2.5]
    d, s, c, t = blend_rrf(gpu_ctx, docs, counts, boosts, 0, 50)   # 0 -> DEFAULT_K = 60
    for q in range(nq):
        wd, ws, wt = oracle.blend_rrf(docs[:, q, :], counts[:, q], boosts, 60, 50)
        assert t[q] == wt and c[q] == len(wd)
        assert np.array_equal(d[q, :c[q]], wd) and np.array_equal(s[q, :c[q]].view(np.uint32), ws.view(np.uint32))


def test_rescore_combine_matches_oracle(gpu_ctx):
    rng = np.random.default_rng(2)
    nq, n = 16, 100
    docs = np.stack([rng.choice(10_000, n, replace=False) for _ in range(nq)]).astype(np.int32)
    scores = np.sort(rng.random((nq, n)).astype(np.float32) * 10)[:, ::-1].copy()
    m = (rng.random((nq, n)) < 0.6).astype(np.uint8)
    s2 = (rng.random((nq, n)) * 5).astype(np.float32)
    d, s = rescore_combine(gpu_ctx, docs, scores, m, s2, 1.0, 2.0)
    for q in range(nq):
        wd, ws = oracle.rescore_combine(docs[q], scores[q], m[q], s2[q], 1.0, 2.0)
        assert np.array_equal(d[q], wd) and np.array_equal(s[q].view(np.uint32), ws.view(np.uint32))
    # QueryRescore constants from the reference tests: qw=1, rw=4 style small ints
    d, s = rescore_combine(gpu_ctx, np.array([[10, 11, 12]], np.int32), np.array([[3., 2., 1.]], np.float32),
                           np.array([[1, 0, 1]], np.uint8), np.array([[1., 0., 4.]], np.float32), 1.0, 2.0)
    assert list(d[0]) == [12, 10, 11] and list(s[0]) == [9.0, 5.0, 2.0]


def test_cross_shard_merge_on_device(gpu_ctx):
    import torch
    rng = np.random.default_rng(3)
    n_lists, nq, k = 4, 33, 50
    scores = np.sort(rng.integers(0, 40, (n_lists, nq, k)).astype(np.float32) / 4, axis=2)[:, :, ::-1].copy()   # many ties
    docs = np.stack([np.stack([np.sort(rng.choice(100_000, k, replace=False)) for _ in range(nq)]) + l * 100_000
                     for l in range(n_lists)]).astype(np.int32)
    # make each list properly ordered (score desc, doc asc)
    for l in range(n_lists):
        for q in range(nq):
            o = np.lexsort((docs[l, q], -scores[l, q]))
            docs[l, q], scores[l, q] = docs[l, q][o], scores[l, q][o]
    counts = rng.integers(0, k + 1, (n_lists, nq)).astype(np.int32)
    dev = torch.device("cuda", 0)
    td, ts, tc = (torch.from_numpy(x).to(dev) for x in (docs, scores, counts))
    od = torch.zeros(nq * k, dtype=torch.int32, device=dev)
    os_ = torch.zeros(nq * k, dtype=torch.float32, device=dev)
    oc = torch.zeros(nq, dtype=torch.int32, device=dev)
    _native.check(_native.gpu_lib().nrtgpu_merge_topk_device(gpu_ctx.handle, n_lists, nq, k, td.data_ptr(), ts.data_ptr(),
                                                             tc.data_ptr(), od.data_ptr(), os_.data_ptr(), oc.data_ptr(),
                                                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    wd, ws, wc = oracle.merge_topk(docs, scores, counts, k)
    gc = oc.cpu().numpy()
    assert np.array_equal(gc, wc)
    gd, gs = od.cpu().numpy().reshape(nq, k), os_.cpu().numpy().reshape(nq, k)
    for q in range(nq):
        assert np.array_equal(gd[q, :gc[q]], wd[q, :gc[q]]) and np.array_equal(gs[q, :gc[q]], ws[q, :gc[q]])


def test_hybrid_pipeline_equals_oracle_pipeline(gpu_ctx):
    """text retriever + kNN retriever + weighted RRF, stage by stage through the C ABI (SearchHandler.executeMultiRetriever
    :528-667 shape): the blended page must equal the oracle pipeline's."""
    from nrtsearch_b200 import index as ix
    from nrtsearch_b200.search import (BooleanQuery, GpuIndex, GpuIndexSearcher, Occur, RelevanceCollector, TermQuery,
                                       compile_queries)
    n, dims, nq, k = 30_000, 64, 24, 50
    sh = ix.synth_text_shard(n, 3_000)
    sh.vectors = ix.synth_vectors(n, dims)
    sh.vec_similarity = ix.SIM_COSINE
    terms = ix.synth_query_terms(nq, 3, 3_000, log10_lo=0.3, log10_hi=3.0)
    qs = [BooleanQuery().add(TermQuery(int(t[0])), Occur.SHOULD).add(TermQuery(int(t[1])), Occur.SHOULD)
          .add(TermQuery(int(t[2])), Occur.SHOULD) for t in terms]
    qv = ix.synth_vectors(nq, dims, seed=ix.SEED_VQUERIES)
    gix = GpuIndex(gpu_ctx, sh)
    s = GpuIndexSearcher(gix)
    t = s.search_batch(qs, RelevanceCollector(k, 2**31 - 1))
    kd, ks, kc = s.knn(qv, k)
    bd, bs, bc, bt = blend_rrf(gpu_ctx, np.stack([t.docs, kd]), np.stack([t.counts, kc]), [1.0, 2.0], 60, k)
    gix.close()
    od, os_, oc, _, _ = oracle.search_compiled(oracle.OracleIndex(sh), *compile_queries(qs), k)
    wkd, wks, wkc = oracle.knn_exact(sh.vectors, ix.SIM_COSINE, qv, k)
    assert np.array_equal(t.docs, od) and np.array_equal(kd, wkd)
    for q in range(nq):
        wd, ws, wt = oracle.blend_rrf(np.stack([od[q], wkd[q]]), [oc[q], wkc[q]], [1.0, 2.0], 60, k)
        assert bt[q] == wt and np.array_equal(bd[q, :bc[q]], wd) and np.array_equal(bs[q, :bc[q]].view(np.uint32), ws.view(np.uint32))


@pytest.mark.parametrize("mode,code", [("max", 1), ("sum", 2), ("avg", 3)])
def test_score_order_blender(gpu_ctx, mode, code):
    """WeightedScoreOrderBlenderOperation.java:50-73 / WeightedScoreDoc.java:57-77: score * boost per retriever, combined by
    MAX / SUM / running AVG in retriever order, float arithmetic; bit-exact vs the oracle (ties by doc asc, as for RRF)."""
    from nrtsearch_b200.search import blend_scores
    rng = np.random.default_rng(71)
    R, nq, top_in, top_out = 3, 40, 60, 50
    docs = np.stack([np.stack([rng.choice(500, size=top_in, replace=False) for _ in range(nq)]) for _ in range(R)]).astype(np.int32)
    scores = np.sort(rng.random((R, nq, top_in)).astype(np.float32) * 10, axis=2)[:, :, ::-1].copy()
    counts = rng.integers(0, top_in + 1, size=(R, nq)).astype(np.int32)
    boosts = np.array([1.0, 0.35, 2.5], np.float32)
    bd, bs, bc, bt = blend_scores(gpu_ctx, mode, docs, scores, counts, boosts, top_out)
    for q in range(nq):
        wd, ws, wt = oracle.blend_scores(code, docs[:, q], scores[:, q], counts[:, q], boosts, top_out)
        assert bt[q] == wt and bc[q] == len(wd)
        assert np.array_equal(bd[q, :bc[q]], wd) and np.array_equal(bs[q, :bc[q]].view(np.uint32), ws.view(np.uint32))
