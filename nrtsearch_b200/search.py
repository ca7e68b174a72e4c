"""Host-side mirror of the reference's query-execution interface for the hot path.

Names and argument meaning follow the Lucene/nrtsearch API at the three call sites the engine replaces
 so parity tests read like the reference's own:

  IndexSearcher.search(Query, CollectorManager)   src/main/java/com/yelp/nrtsearch/server/handler/SearchHandler.java:1412
  BooleanQuery / TermQuery / RangeQuery / BoostQuery / MatchAllDocsQuery
                                                  src/main/java/com/yelp/nrtsearch/server/query/QueryNodeMapper.java:257-283
  RelevanceCollector (numHitsToCollect, totalHitsThreshold, searchAfter)
                                                  src/main/java/com/yelp/nrtsearch/server/search/collectors/RelevanceCollector.java:42-69
  KnnQuery / ExactVectorQuery                     src/main/java/com/yelp/nrtsearch/server/search/KnnUtils.java:47-66

All execution happens in libnrtgpu.so (CUDA); this module only marshals.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from ._native import (Aggregation as CAgg, AggregationResult as CAggResult, Clause, CollectionTimeoutException, NrtGpuError,
                      NrtGpuUnsupported, Query as CQuery, SearchLimits, Sort as CSort, check)
from .index import HostShard, PinnedDesc

TOTAL_HITS_THRESHOLD = 1000  # SearchRequestProcessor.TOTAL_HITS_THRESHOLD (:102)
INT_MAX = 2**31 - 1


class Occur(enum.IntEnum):
    SHOULD = 0
    MUST = 1
    FILTER = 2
    MUST_NOT = 3


@dataclass(frozen=True)
class TermQuery:
    term: int  # the adaptor's dense (field, term) id


@dataclass(frozen=True)
class RangeQuery:
    """IndexOrDocValuesQuery(PointRangeQuery, SortedNumericDocValuesRangeQuery): inclusive bounds in the
    sortable-long domain (IntFieldDef.getRangeQuery :124-158 folds exclusive bounds with +-1)."""
    column: int
    lower: int = -(2**63)
    upper: int = 2**63 - 1


@dataclass(frozen=True)
class MatchAllDocsQuery:
    pass


@dataclass(frozen=True)
class BoostQuery:
    query: object
    boost: float


@dataclass(frozen=True)
class BooleanClause:
    query: object
    occur: Occur


@dataclass
class BooleanQuery:
    clauses: List[BooleanClause] = field(default_factory=list)
    minimum_number_should_match: int = 0

    def add(self, query, occur: Occur) -> "BooleanQuery":
        self.clauses.append(BooleanClause(query, Occur(occur)))
        return self


def boolean_query_from_proto(clauses: Sequence[Tuple[object, Occur]], minimum_number_should_match: int = 0) -> BooleanQuery:
    """QueryNodeMapper.getBooleanQuery (:257-283): empty => MatchAllDocs MUST; all MUST_NOT => add MatchAllDocs FILTER."""
    bq = BooleanQuery(minimum_number_should_match=minimum_number_should_match)
    if not clauses:
        return bq.add(MatchAllDocsQuery(), Occur.MUST)
    all_must_not = True
    for q, occ in clauses:
        bq.add(q, occ)
        if occ != Occur.MUST_NOT:
            all_must_not = False
    if all_must_not:
        bq.add(MatchAllDocsQuery(), Occur.FILTER)
    return bq


@dataclass
class ScoreDoc:
    doc: int
    score: float


class Relation(enum.IntEnum):
    EQUAL_TO = 0
    GREATER_THAN_OR_EQUAL_TO = 1


@dataclass
class TotalHits:
    value: int
    relation: Relation


@dataclass
class TopDocs:
    total_hits: TotalHits
    score_docs: List[ScoreDoc]


@dataclass
class RelevanceCollector:
    """DocCollector config (CollectorCreatorContext.java:36-53, DocCollector.java wrappers): numHitsToCollect,
    totalHitsThreshold, searchAfter; timeoutSec / disallowPartialResults (SearchCutoffWrapper.java:164-202);
    terminateAfter / terminateAfterMaxRecallCount (TerminateAfterWrapper.java:85-162)."""
    num_hits_to_collect: int
    total_hits_threshold: int = TOTAL_HITS_THRESHOLD
    search_after: Optional[ScoreDoc] = None
    timeout_sec: float = 0.0
    elapsed_sec: float = 0.0
    disallow_partial_results: bool = False
    terminate_after: int = 0
    terminate_after_max_recall_count: int = 0

    def limits(self) -> Optional[SearchLimits]:
        if self.timeout_sec <= 0 and self.terminate_after <= 0:
            return None
        return SearchLimits(self.timeout_sec, self.elapsed_sec, 1 if self.disallow_partial_results else 0,
                            self.terminate_after, self.terminate_after_max_recall_count)


def float_to_sortable_int(f: float) -> int:
    """NumericUtils.floatToSortableInt: the order-preserving int of a float (how FloatFieldDef stores doc values)."""
    b = int(np.float32(f).view(np.int32))
    return b ^ ((b >> 31) & 0x7fffffff)


def double_to_sortable_long(d: float) -> int:
    """NumericUtils.doubleToSortableLong."""
    b = int(np.float64(d).view(np.int64))
    return b ^ ((b >> 63) & 0x7fffffffffffffff)


@dataclass(frozen=True)
class SortType:
    """SortType of the search request for ONE sort field (SortParser.parseSort :54-95): a numeric doc-value column, or
    "docid". field_type picks the missing value exactly as the reference's FieldDefs do (IntFieldDef.java:103,
    LongFieldDef.java:103, FloatFieldDef.java:105, DoubleFieldDef.java:105): MAX / +Infinity when missing_last, else
    MIN / -Infinity -- irrespective of reverse."""
    field: object            # column id, or "docid"
    reverse: bool = False
    missing_last: bool = False
    field_type: str = "long"   # int | long | float | double

    def missing_value(self) -> int:
        hi = self.missing_last
        if self.field_type == "int":
            return 2**31 - 1 if hi else -(2**31)
        if self.field_type == "long":
            return 2**63 - 1 if hi else -(2**63)
        if self.field_type == "float":
            return float_to_sortable_int(float("inf") if hi else float("-inf"))
        if self.field_type == "double":
            return double_to_sortable_long(float("inf") if hi else float("-inf"))
        raise ValueError(f"field type {self.field_type} does not support sorting")


@dataclass
class SortFieldCollector:
    """SortFieldCollector.java:44-105: numHitsToCollect + the query's Sort; searchAfter is a FieldDoc (value, doc)."""
    num_hits_to_collect: int
    sort: SortType = None
    timeout_sec: float = 0.0
    terminate_after: int = 0


@dataclass(frozen=True)
class TermsCollector:
    """TermsCollector over a numeric doc-value field ({Int,Long,Float,Double}TermsCollectorManager): `size` buckets ordered
    by count (BucketOrder COUNT, desc by default)."""
    column: int
    size: int
    order_desc: bool = True
    field_type: str = "long"


@dataclass(frozen=True)
class MinCollector:
    column: int
    field_type: str = "long"


@dataclass(frozen=True)
class MaxCollector:
    column: int
    field_type: str = "long"


@dataclass(frozen=True)
class SumCollector:
    column: int
    field_type: str = "long"


_VALUE_TYPE = {"int": 0, "long": 0, "float": 1, "double": 2}


@dataclass
class FieldDoc:
    doc: int
    value: int   # fields[0], sortable-long domain


@dataclass
class SortedResult:
    docs: np.ndarray          # int32 [nq, k]
    sort_values: np.ndarray   # int64 [nq, k] FieldDoc.fields[0] of every hit
    counts: np.ndarray
    total_hits: np.ndarray
    relation: np.ndarray


def _f32(x: float) -> np.float32:
    return np.float32(x)


def _flatten(q, boost: np.float32, out: list, occur: Occur) -> None:
    """One clause of the flat BooleanQuery; BoostQuery boosts multiply outermost-first in float
    (BoostQuery.createWeight passes boost * this.boost down)."""
    while isinstance(q, BoostQuery):
        if q.boost < 0:
            raise ValueError("Boost must be a positive number")  # QueryNodeMapper.java:127
        boost = _f32(boost * _f32(q.boost))
        q = q.query
    if isinstance(q, TermQuery):
        out.append((int(occur), 0, int(q.term), float(boost), 0, 0))
    elif isinstance(q, RangeQuery):
        out.append((int(occur), 1, int(q.column), float(boost), int(q.lower), int(q.upper)))
    elif isinstance(q, MatchAllDocsQuery):
        out.append((int(occur), 2, 0, float(boost), 0, 0))
    else:
        raise NrtGpuUnsupported(3, f"query node {type(q).__name__} is outside the GPU path")


def compile_queries(queries: Sequence[object], search_after: Optional[Sequence[Optional[ScoreDoc]]] = None):
    """Query trees -> (Clause[], Query[]) for nrtgpu_search_bool. A bare leaf is a single MUST clause
    (Lucene rewrites a one-clause BooleanQuery to its clause; scores are identical)."""
    flat, qs = [], []
    for i, q in enumerate(queries):
        boost = _f32(1.0)
        while isinstance(q, BoostQuery):
            if q.boost < 0:
                raise ValueError("Boost must be a positive number")
            boost = _f32(boost * _f32(q.boost))
            q = q.query
        begin = len(flat)
        msm = 0
        if isinstance(q, BooleanQuery):
            msm = q.minimum_number_should_match
            for cl in q.clauses:
                if isinstance(cl.query, BooleanQuery):
                    raise NrtGpuUnsupported(3, "nested BooleanQuery is outside the GPU path")
                _flatten(cl.query, boost, flat, cl.occur)
        else:
            _flatten(q, boost, flat, Occur.MUST)
        after = search_after[i] if search_after is not None else None
        qs.append((begin, len(flat), msm, 1 if after is not None else 0,
                   after.doc if after is not None else 0, after.score if after is not None else 0.0))
    carr = (Clause * max(len(flat), 1))()
    for i, (occ, kind, id_, b, lo, hi) in enumerate(flat):
        carr[i] = Clause(occ, kind, id_, b, lo, hi)
    qarr = (CQuery * max(len(qs), 1))()
    for i, t in enumerate(qs):
        qarr[i] = CQuery(*t)
    return carr, len(flat), qarr, len(qs)


class GpuContext:
    def __init__(self, device: int = 0):
        self._lib = _native.gpu_lib()
        h = C.c_void_p()
        check(self._lib.nrtgpu_init(device, C.byref(h)))
        self.handle = h
        self.device = device

    def close(self):
        if self.handle:
            self._lib.nrtgpu_shutdown(self.handle)
            self.handle = None


class GpuIndex:
    """Device image of one shard at one reader version (ShardSearcherFactory.newSearcher hook)."""

    def __init__(self, ctx: GpuContext, shard: HostShard):
        self._lib = _native.gpu_lib()
        self.ctx = ctx
        pinned = PinnedDesc(shard)
        h = C.c_void_p()
        check(self._lib.nrtgpu_index_build(ctx.handle, C.byref(pinned.desc), C.byref(h)))
        self.handle = h
        self.n_docs, self.doc_base = shard.n_docs, shard.doc_base

    def set_live_docs(self, live_docs: Optional[np.ndarray]):
        """Deletes of a new reader version (LeafReader.getLiveDocs): refreshed in place, no image rebuild."""
        lv = None if live_docs is None else np.ascontiguousarray(live_docs, np.uint8)
        check(self._lib.nrtgpu_index_set_live_docs(self.handle, None if lv is None else lv.ctypes.data))

    def update_stats(self, term_df: Optional[np.ndarray], field_doc_count: Sequence[int], field_sum_ttf: Sequence[int]):
        """Index-wide BM25 statistics changed (a leaf was added elsewhere in the shard): idf inputs, length caches, impacts."""
        df = None if term_df is None else np.ascontiguousarray(term_df, np.int64)
        dc = np.ascontiguousarray(field_doc_count, np.int64)
        tt = np.ascontiguousarray(field_sum_ttf, np.int64)
        check(self._lib.nrtgpu_index_update_stats(self.handle, None if df is None else df.ctypes.data, dc.ctypes.data, tt.ctypes.data))

    @property
    def device_bytes(self) -> int:
        return int(self._lib.nrtgpu_index_device_bytes(self.handle))

    def close(self):
        if self.handle:
            self._lib.nrtgpu_index_close(self.handle)
            self.handle = None


@dataclass
class BatchResult:
    docs: np.ndarray      # int32 [nq, k]
    scores: np.ndarray    # float32 [nq, k]
    counts: np.ndarray    # int32 [nq]
    total_hits: np.ndarray  # int64 [nq]
    relation: np.ndarray  # uint8 [nq]
    hit_timeout: Optional[np.ndarray] = None        # uint8 [nq] (SearchResponse.hitTimeout, per query of the batch)
    terminated_early: Optional[np.ndarray] = None   # uint8 [nq] (SearchResponse.terminatedEarly)

    def top_docs(self, i: int) -> TopDocs:
        n = int(self.counts[i])
        return TopDocs(TotalHits(int(self.total_hits[i]), Relation(int(self.relation[i]))),
                       [ScoreDoc(int(d), float(s)) for d, s in zip(self.docs[i, :n], self.scores[i, :n])])


class PreparedBatch:
    """nrtgpu_batch: compiled batch resident on the device (launch many times, inputs stay in HBM)."""

    def __init__(self, index: GpuIndex, carr, ncl, qarr, nq, top_k, threshold, flags=0):
        self._lib = _native.gpu_lib()
        self.index, self.nq, self.top_k = index, nq, top_k
        h = C.c_void_p()
        check(self._lib.nrtgpu_batch_prepare(index.handle, carr, ncl, qarr, nq, top_k, threshold, flags, C.byref(h)))
        self.handle = h

    def run(self, stream: int = 0):
        check(self._lib.nrtgpu_batch_run(self.handle, C.c_void_p(stream)))

    def fetch(self, stream: int = 0, out: Optional[BatchResult] = None) -> BatchResult:
        if out is None:
            out = BatchResult(np.zeros((self.nq, self.top_k), np.int32), np.zeros((self.nq, self.top_k), np.float32),
                              np.zeros(self.nq, np.int32), np.zeros(self.nq, np.int64), np.zeros(self.nq, np.uint8))
        check(self._lib.nrtgpu_batch_fetch(self.handle, C.c_void_p(stream), out.docs.ctypes.data, out.scores.ctypes.data,
                                           out.counts.ctypes.data, out.total_hits.ctypes.data, out.relation.ctypes.data))
        return out

    def stats(self):
        a, l, w = C.c_int64(), C.c_int32(), C.c_int64()
        check(self._lib.nrtgpu_batch_stats(self.handle, C.byref(a), C.byref(l), C.byref(w)))
        return {"alg_postings": a.value, "launches_per_run": l.value, "work_items": w.value}

    def stage_ms(self, stage: int) -> float:
        ms = C.c_float()
        check(self._lib.nrtgpu_batch_stage_ms(self.handle, stage, C.byref(ms)))
        return ms.value

    def reset_timing(self):
        check(self._lib.nrtgpu_batch_reset_timing(self.handle))

    def bind_output(self, d_docs: int, d_scores: int, d_counts: int):
        check(self._lib.nrtgpu_batch_bind_output(self.handle, C.c_void_p(d_docs), C.c_void_p(d_scores), C.c_void_p(d_counts)))

    def bind_packed(self, d_record: int):
        """Results of subsequent runs go into one packed DEVICE record (nrtgpu_batch_bind_packed): the buffer a multi-GPU
        step all-gathers."""
        check(self._lib.nrtgpu_batch_bind_packed(self.handle, C.c_void_p(d_record)))

    def device_results(self):
        d, s, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self._lib.nrtgpu_batch_device_results(self.handle, C.byref(d), C.byref(s), C.byref(c)))
        return d.value, s.value, c.value

    def close(self):
        if self.handle:
            self._lib.nrtgpu_batch_free(self.handle)
            self.handle = None


class GpuIndexSearcher:
    """Batched stand-in for MyIndexSearcher.search(Query, CollectorManager)."""

    def __init__(self, index: GpuIndex):
        self._lib = _native.gpu_lib()
        self.index = index

    def prepare(self, queries: Sequence[object], collector: RelevanceCollector,
                search_after: Optional[Sequence[Optional[ScoreDoc]]] = None, flags: int = 0) -> PreparedBatch:
        if search_after is None and collector.search_after is not None:
            search_after = [collector.search_after] * len(queries)
        carr, ncl, qarr, nq = compile_queries(queries, search_after)
        return PreparedBatch(self.index, carr, ncl, qarr, nq, collector.num_hits_to_collect,
                             collector.total_hits_threshold, flags)

    def search_batch(self, queries: Sequence[object], collector: RelevanceCollector,
                     search_after: Optional[Sequence[Optional[ScoreDoc]]] = None, stream: int = 0) -> BatchResult:
        """One call through the C ABI with HOST buffers (nrtgpu_search_bool)."""
        if search_after is None and collector.search_after is not None:
            search_after = [collector.search_after] * len(queries)
        carr, ncl, qarr, nq = compile_queries(queries, search_after)
        k = collector.num_hits_to_collect
        out = BatchResult(np.zeros((nq, max(k, 1)), np.int32), np.zeros((nq, max(k, 1)), np.float32),
                          np.zeros(nq, np.int32), np.zeros(nq, np.int64), np.zeros(nq, np.uint8),
                          np.zeros(nq, np.uint8), np.zeros(nq, np.uint8))
        lim = collector.limits()
        check(self._lib.nrtgpu_search_bool_ex(self.index.handle, carr, ncl, qarr, nq, k, collector.total_hits_threshold, 0,
                                              None if lim is None else C.byref(lim), C.c_void_p(stream), out.docs.ctypes.data,
                                              out.scores.ctypes.data, out.counts.ctypes.data, out.total_hits.ctypes.data,
                                              out.relation.ctypes.data, out.hit_timeout.ctypes.data,
                                              out.terminated_early.ctypes.data))
        return out

    def knn_query(self, queries: np.ndarray, knn: KnnQuery, sim: int, boosts: Optional[np.ndarray] = None,
                  filter_docs: Optional[np.ndarray] = None, stream: int = 0):
        """KnnUtils.resolveKnnQueryAndBoost (:47-66) for a batch of query vectors under one KnnQuery configuration."""
        knn.validate()
        docs, scores, counts = self.knn(queries, knn.k, boosts, filter_docs, stream)
        if knn.similarity_threshold is not None:   # MinThresholdQuery: drop hits scoring below the threshold's score
            th = similarity_to_score(knn.similarity_threshold, sim, queries.shape[1])
            for q in range(len(counts)):
                b = np.float32(1.0) if boosts is None else np.float32(boosts[q])
                keep = scores[q, :counts[q]] >= th * b
                n = int(keep.sum())
                docs[q, :n] = docs[q, :counts[q]][keep]; scores[q, :n] = scores[q, :counts[q]][keep]
                docs[q, n:] = 0; scores[q, n:] = 0; counts[q] = n
        return docs, scores, counts

    def search_sorted(self, queries: Sequence[object], collector: SortFieldCollector,
                      search_after: Optional[Sequence[Optional[FieldDoc]]] = None, stream: int = 0) -> SortedResult:
        """IndexSearcher.search(query, TopFieldCollectorManager(sort, numHits, after, threshold)) for a batch."""
        st = collector.sort
        after_sd = None if search_after is None else [None if a is None else ScoreDoc(a.doc, 0.0) for a in search_after]
        carr, ncl, qarr, nq = compile_queries(queries, after_sd)
        k = collector.num_hits_to_collect
        out = SortedResult(np.zeros((nq, k), np.int32), np.zeros((nq, k), np.int64), np.zeros(nq, np.int32), np.zeros(nq, np.int64),
                           np.zeros(nq, np.uint8))
        av = None
        if search_after is not None:
            av = np.array([0 if a is None else a.value for a in search_after], np.int64)
        docid = st.field == "docid"
        cs = CSort(2 if docid else 1, 0 if docid else int(st.field), 1 if st.reverse else 0, 0, 0 if docid else st.missing_value(),
                   None if av is None else av.ctypes.data)
        lim = None
        if collector.timeout_sec > 0 or collector.terminate_after > 0:
            lim = SearchLimits(collector.timeout_sec, 0.0, 0, collector.terminate_after, 0)
        check(self._lib.nrtgpu_search_sorted(self.index.handle, carr, ncl, qarr, nq, k, 0, C.byref(cs), None if lim is None else C.byref(lim),
                                             C.c_void_p(stream), out.docs.ctypes.data, out.sort_values.ctypes.data, out.counts.ctypes.data,
                                             out.total_hits.ctypes.data, out.relation.ctypes.data, None, None))
        return out

    def search_with_collectors(self, queries: Sequence[object], collector: RelevanceCollector, additional: Sequence[object],
                               stream: int = 0):
        """IndexSearcher.search with additional collectors (SearchCollectorManager fan-out): returns (BatchResult, results)
        where results[i] is a float64 [nq] array (min / max / sum) or a dict of bucket arrays (terms)."""
        carr, ncl, qarr, nq = compile_queries(queries)
        k = collector.num_hits_to_collect
        out = BatchResult(np.zeros((nq, k), np.int32), np.zeros((nq, k), np.float32), np.zeros(nq, np.int32), np.zeros(nq, np.int64),
                          np.zeros(nq, np.uint8))
        aggs = (CAgg * len(additional))()
        res = (CAggResult * len(additional))()
        outs = []
        for i, a in enumerate(additional):
            vt = _VALUE_TYPE[a.field_type]
            if isinstance(a, TermsCollector):
                aggs[i] = CAgg(1, a.column, vt, a.size, 1 if a.order_desc else 0, 0)
                o = {"keys": np.zeros((nq, a.size), np.int64), "counts": np.zeros((nq, a.size), np.int32), "n": np.zeros(nq, np.int32),
                     "total_buckets": np.zeros(nq, np.int32), "other_counts": np.zeros(nq, np.int64)}
                res[i] = CAggResult(None, o["keys"].ctypes.data, o["counts"].ctypes.data, o["n"].ctypes.data,
                                    o["total_buckets"].ctypes.data, o["other_counts"].ctypes.data)
            else:
                kind = 2 if isinstance(a, MinCollector) else 3 if isinstance(a, MaxCollector) else 4
                aggs[i] = CAgg(kind, a.column, vt, 0, 0, 0)
                o = np.zeros(nq, np.float64)
                res[i] = CAggResult(o.ctypes.data, None, None, None, None, None)
            outs.append(o)
        check(self._lib.nrtgpu_search_bool_aggs(self.index.handle, carr, ncl, qarr, nq, k, 0, aggs, len(additional), res, C.c_void_p(stream),
                                                out.docs.ctypes.data, out.scores.ctypes.data, out.counts.ctypes.data,
                                                out.total_hits.ctypes.data))
        return out, outs

    def score_docs(self, queries: Sequence[object], docs: np.ndarray, counts: Optional[np.ndarray] = None, stream: int = 0):
        """Second pass of QueryRescorer: query q on its own hit list -> (matches uint8 [nq, n], scores float32 [nq, n])."""
        carr, ncl, qarr, nq = compile_queries(queries)
        d = np.ascontiguousarray(docs, np.int32)
        cn = None if counts is None else np.ascontiguousarray(counts, np.int32)
        m, s = np.zeros(d.shape, np.uint8), np.zeros(d.shape, np.float32)
        check(self._lib.nrtgpu_score_docs(self.index.handle, carr, ncl, qarr, nq, d.shape[1], d.ctypes.data, None if cn is None else cn.ctypes.data,
                                          C.c_void_p(stream), m.ctypes.data, s.ctypes.data))
        return m, s

    def rescore_query(self, queries: Sequence[object], docs: np.ndarray, scores: np.ndarray, counts: np.ndarray, window: int,
                      query_weight: float, rescore_weight: float, stream: int = 0):
        """QueryRescore (QueryRescore.java:39-57) end to end on the device: returns docs, scores, counts of the rescored lists."""
        carr, ncl, qarr, nq = compile_queries(queries)
        d = np.ascontiguousarray(docs, np.int32).copy()
        s = np.ascontiguousarray(scores, np.float32).copy()
        cn = np.ascontiguousarray(counts, np.int32)
        oc = np.zeros(nq, np.int32)
        check(self._lib.nrtgpu_rescore_query(self.index.handle, carr, ncl, qarr, nq, d.shape[1], cn.ctypes.data, window, query_weight,
                                             rescore_weight, C.c_void_p(stream), d.ctypes.data, s.ctypes.data, oc.ctypes.data))
        return d, s, oc

    def fetch_columns(self, columns: Sequence[int], docs: np.ndarray, stream: int = 0):
        """Fetch phase on doc-value columns: values int64 [n_cols, n], has uint8 [n_cols, n] for the hits `docs`."""
        cols = np.ascontiguousarray(columns, np.int32)
        d = np.ascontiguousarray(docs, np.int32).reshape(-1)
        vals, has = np.zeros((len(cols), len(d)), np.int64), np.zeros((len(cols), len(d)), np.uint8)
        check(self._lib.nrtgpu_fetch_columns(self.index.handle, cols.ctypes.data, len(cols), d.ctypes.data, len(d), C.c_void_p(stream),
                                             vals.ctypes.data, has.ctypes.data))
        return vals, has

    def search(self, query, collector: RelevanceCollector) -> TopDocs:
        return self.search_batch([query], collector).top_docs(0)

    def knn(self, queries: np.ndarray, k: int, boosts: Optional[np.ndarray] = None,
            filter_docs: Optional[np.ndarray] = None, stream: int = 0):
        """Exact kNN (KnnUtils.resolveKnnQueryAndBoost with ExactVectorQuery semantics): returns docs, scores, counts."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        docs = np.zeros((nq, k), np.int32)
        scores = np.zeros((nq, k), np.float32)
        counts = np.zeros(nq, np.int32)
        b = None if boosts is None else np.ascontiguousarray(boosts, dtype=np.float32)
        f = None if filter_docs is None else np.ascontiguousarray(filter_docs, dtype=np.uint8)
        check(self._lib.nrtgpu_search_knn(self.index.handle, q.ctypes.data, nq, k,
                                          None if b is None else b.ctypes.data, None if f is None else f.ctypes.data,
                                          C.c_void_p(stream), docs.ctypes.data, scores.ctypes.data, counts.ctypes.data))
        return docs, scores, counts


class GpuLeafSearcher:
    """IndexSearcher over the leaf images of one reader version (nrtgpu_searcher_*): every leaf runs the batch, pages are
    merged on the device (TopDocs.merge). A new NRT reader version = the old leaves' images + images of the new leaves."""

    def __init__(self, ctx: GpuContext, leaves: Sequence[GpuIndex]):
        self._lib = _native.gpu_lib()
        arr = (C.c_void_p * len(leaves))(*[l.handle for l in leaves])
        h = C.c_void_p()
        check(self._lib.nrtgpu_searcher_create(ctx.handle, arr, len(leaves), C.byref(h)))
        self.handle, self.leaves = h, list(leaves)

    def search_batch(self, queries: Sequence[object], collector: RelevanceCollector, stream: int = 0) -> BatchResult:
        carr, ncl, qarr, nq = compile_queries(queries)
        k = collector.num_hits_to_collect
        out = BatchResult(np.zeros((nq, k), np.int32), np.zeros((nq, k), np.float32), np.zeros(nq, np.int32), np.zeros(nq, np.int64),
                          np.zeros(nq, np.uint8))
        lim = collector.limits()
        check(self._lib.nrtgpu_searcher_search_bool(self.handle, carr, ncl, qarr, nq, k, collector.total_hits_threshold, 0,
                                                    None if lim is None else C.byref(lim), C.c_void_p(stream), out.docs.ctypes.data,
                                                    out.scores.ctypes.data, out.counts.ctypes.data, out.total_hits.ctypes.data,
                                                    out.relation.ctypes.data))
        return out

    def close(self):
        if self.handle:
            self._lib.nrtgpu_searcher_close(self.handle)
            self.handle = None


class GpuBatcher:
    """Request micro-batcher (nrtgpu_batcher_*): SearchHandler threads submit ONE query each and block; a native worker
    thread turns the waiting requests into batched nrtgpu_search_bool calls. The Java adaptor's counterpart is
    jni/java/.../GpuIndexSearcher.java (GpuBatcher.forIndex)."""

    def __init__(self, index: GpuIndex, max_batch: int = 256, max_wait_us: int = 200):
        self._lib = _native.gpu_lib()
        h = C.c_void_p()
        check(self._lib.nrtgpu_batcher_create(index.handle, max_batch, max_wait_us, C.byref(h)))
        self.handle = h

    def submit(self, query, collector: RelevanceCollector):
        """Blocking: returns (TopDocs, Diagnostics) of the one query."""
        carr, ncl, qarr, _ = compile_queries([query])
        k = collector.num_hits_to_collect
        docs, scores = np.zeros(k, np.int32), np.zeros(k, np.float32)
        cnt, tot, rel = C.c_int32(), C.c_int64(), C.c_uint8()
        diag = _native.Diagnostics()
        check(self._lib.nrtgpu_batcher_submit(self.handle, carr, ncl, qarr[0].min_should_match, k, collector.total_hits_threshold,
                                              docs.ctypes.data, scores.ctypes.data, C.byref(cnt), C.byref(tot), C.byref(rel), C.byref(diag)))
        n = cnt.value
        return TopDocs(TotalHits(tot.value, Relation(rel.value)), [ScoreDoc(int(d), float(s)) for d, s in zip(docs[:n], scores[:n])]), diag

    def stats(self):
        a, b = C.c_int64(), C.c_int64()
        check(self._lib.nrtgpu_batcher_stats(self.handle, C.byref(a), C.byref(b)))
        return {"batches": a.value, "requests": b.value}

    def close(self):
        if self.handle:
            self._lib.nrtgpu_batcher_close(self.handle)
            self.handle = None


NUM_CANDIDATES_LIMIT = 10000   # VectorFieldDef.java:74


@dataclass(frozen=True)
class KnnQuery:
    """KnnQuery of the search request as VectorFieldDef.getKnnQuery validates it (VectorFieldDef.java:401-425). The
    reference runs HNSW with a beam of num_candidates PER LEAF and merges the leaves' lists to k
    (NrtKnnFloatVectorQuery.java:43-64); here every leaf / shard is searched EXACTLY, so any num_candidates >= k yields the
    same -- exact -- top k (recall 1.0); it is validated and otherwise unused. similarity_threshold wraps the query in
    MinThresholdQuery(score >= similarityToScore(threshold)) (:590-593)."""
    k: int
    num_candidates: int
    similarity_threshold: Optional[float] = None

    def validate(self):
        if self.k < 1:
            raise ValueError("Vector search k must be >= 1")
        if self.num_candidates < self.k:
            raise ValueError("Vector search numCandidates must be >= k")
        if self.num_candidates > NUM_CANDIDATES_LIMIT:
            raise ValueError(f"Vector search numCandidates > {NUM_CANDIDATES_LIMIT}")


def similarity_to_score(similarity: float, sim: int, dims: int = 0, byte_field: bool = False) -> np.float32:
    """VectorFieldDef.similarityToScore (float :664-673, byte :870-881), float arithmetic."""
    s = np.float32(similarity)
    if sim == 0:
        return np.float32(1.0) / (np.float32(1.0) + s * s)
    if sim == 1 and byte_field:
        return np.float32(0.5) + s / np.float32(dims * (1 << 15))
    if sim in (1, 2):
        return (np.float32(1.0) + s) / np.float32(2.0)
    return np.float32(1.0) / (np.float32(1.0) + np.float32(-1.0) * s) if s < 0 else s + np.float32(1.0)


def normalized_cosine_vectors(vectors: np.ndarray):
    """What a `normalized_cosine` vector field does to its input (VectorFieldDef.java:308-332, 507-513, 568-573, 651-655):
    magnitude = sqrt(float dot(v, v)); v /= magnitude (float); the field is then searched with DOT_PRODUCT and the magnitude is
    kept in the <field>._magnitude float doc value. Returns (unit vectors float32, magnitudes float32)."""
    v = np.ascontiguousarray(vectors, np.float32)
    mag = np.sqrt(np.einsum("ij,ij->i", v, v, dtype=np.float32)).astype(np.float32)
    if (mag == 0).any():
        raise ValueError("Vector magnitude cannot be 0 when using cosine similarity")   # validateVectorForSearch :634-639
    return (v / mag[:, None]).astype(np.float32), mag


def blend_rrf(ctx: GpuContext, docs: np.ndarray, counts: np.ndarray, boosts: Sequence[float], rank_constant: int,
              top_hits: int):
    """BlenderOperation.blend with the weighted-RRF operation (BlenderOperation.java:76-87) for nq queries:
    docs [R, nq, top_in], counts [R, nq] -> (docs [nq, top_hits], scores, counts, total)."""
    d = np.ascontiguousarray(docs, np.int32)
    c = np.ascontiguousarray(counts, np.int32)
    b = np.ascontiguousarray(boosts, np.float32)
    R, nq, top_in = d.shape
    od, os_ = np.zeros((nq, top_hits), np.int32), np.zeros((nq, top_hits), np.float32)
    oc, ot = np.zeros(nq, np.int32), np.zeros(nq, np.int32)
    check(_native.gpu_lib().nrtgpu_blend_rrf(ctx.handle, R, nq, top_in, d.ctypes.data, c.ctypes.data, b.ctypes.data,
                                             rank_constant, top_hits, od.ctypes.data, os_.ctypes.data, oc.ctypes.data,
                                             ot.ctypes.data))
    return od, os_, oc, ot


def blend_scores(ctx: GpuContext, mode: str, docs: np.ndarray, scores: np.ndarray, counts: np.ndarray, boosts: Sequence[float],
                 top_hits: int):
    """BlenderOperation.blend with WeightedScoreOrderBlenderOperation (MAX / SUM / AVG of score * boost): docs, scores
    [R, nq, top_in], counts [R, nq] -> (docs [nq, top_hits], scores, counts, total)."""
    d = np.ascontiguousarray(docs, np.int32)
    s = np.ascontiguousarray(scores, np.float32)
    c = np.ascontiguousarray(counts, np.int32)
    b = np.ascontiguousarray(boosts, np.float32)
    R, nq, top_in = d.shape
    od, os_ = np.zeros((nq, top_hits), np.int32), np.zeros((nq, top_hits), np.float32)
    oc, ot = np.zeros(nq, np.int32), np.zeros(nq, np.int32)
    check(_native.gpu_lib().nrtgpu_blend_scores(ctx.handle, {"max": 1, "sum": 2, "avg": 3}[mode.lower()], R, nq, top_in, d.ctypes.data,
                                                s.ctypes.data, c.ctypes.data, b.ctypes.data, top_hits, od.ctypes.data, os_.ctypes.data,
                                                oc.ctypes.data, ot.ctypes.data))
    return od, os_, oc, ot


def rescore_combine(ctx: GpuContext, docs: np.ndarray, scores: np.ndarray, second_matches: np.ndarray,
                    second_scores: np.ndarray, query_weight: float, rescore_weight: float, counts=None):
    """QueryRescore (QueryRescore.java:39-57) for nq hit lists [nq, n_hits]: combined + re-sorted copies."""
    d = np.ascontiguousarray(docs, np.int32).copy()
    s = np.ascontiguousarray(scores, np.float32).copy()
    m = np.ascontiguousarray(second_matches, np.uint8)
    s2 = np.ascontiguousarray(second_scores, np.float32)
    nq, n_hits = d.shape
    cn = None if counts is None else np.ascontiguousarray(counts, np.int32)
    check(_native.gpu_lib().nrtgpu_rescore_combine(ctx.handle, nq, n_hits, None if cn is None else cn.ctypes.data,
                                                   d.ctypes.data, s.ctypes.data, m.ctypes.data, s2.ctypes.data,
                                                   query_weight, rescore_weight))
    return d, s
