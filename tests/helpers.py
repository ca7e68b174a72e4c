"""Test helpers: tiny hand-typed corpora -> HostShard, result comparison."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

import oracle
from nrtsearch_b200.index import HostShard, TextField


def shard_from_token_docs(field_docs: Sequence[Sequence[Sequence[str]]], omit_norms: Sequence[bool] = None,
                          columns: Optional[List[np.ndarray]] = None, live_docs=None):
    """field_docs[f][d] = token list of doc d in text field f (empty list = doc lacks the field).
    Returns (HostShard, vocab) with vocab[(f, token)] = term id. Statistics as Lucene computes them:
    docCount = docs having the field, sumTotalTermFreq = total tokens, norm = intToByte4(length)."""
    n_fields = len(field_docs)
    n_docs = len(field_docs[0])
    vocab: Dict = {}
    postings: Dict[int, Dict[int, int]] = {}
    term_field: List[int] = []
    fields = []
    for f in range(n_fields):
        norms = np.zeros(n_docs, np.uint8)
        doc_count = ttf = 0
        for d, toks in enumerate(field_docs[f]):
            if not toks:
                continue
            doc_count += 1
            ttf += len(toks)
            norms[d] = oracle.int_to_byte4(len(toks))
            for t in toks:
                tid = vocab.setdefault((f, t), len(vocab))
                if tid == len(term_field):
                    term_field.append(f)
                postings.setdefault(tid, {})
                postings[tid][d] = postings[tid].get(d, 0) + 1
        fields.append(TextField(None if (omit_norms and omit_norms[f]) else norms, doc_count, ttf))
    nt = len(vocab)
    off = np.zeros(nt + 1, np.int64)
    docs, freqs = [], []
    for t in range(nt):
        ds = sorted(postings[t])
        docs += ds
        freqs += [postings[t][d] for d in ds]
        off[t + 1] = len(docs)
    sh = HostShard(n_docs=n_docs, doc_base=0, term_off=off, post_docs=np.array(docs, np.int32),
                   post_freqs=np.array(freqs, np.int32), fields=fields, term_field=np.array(term_field, np.int32),
                   columns=columns or [], column_has=[None] * len(columns or []), live_docs=live_docs)
    return sh, vocab


def assert_same_hits(got, want, check_total=True, what=""):
    """Parity spec: counts, doc sequence, bit-identical scores, totalHits when EQUAL_TO."""
    gd, gs, gc, gt, gr = got
    wd, ws, wc, wt, wr = want
    assert np.array_equal(gc, wc), f"{what} counts differ: {gc[:8]} vs {wc[:8]}"
    for q in range(len(gc)):
        n = int(gc[q])
        if not np.array_equal(gd[q, :n], wd[q, :n]):
            bad = np.nonzero(gd[q, :n] != wd[q, :n])[0][0]
            raise AssertionError(f"{what} query {q}: doc mismatch at rank {bad}: {gd[q, bad]} ({gs[q, bad]!r}) vs "
                                 f"{wd[q, bad]} ({ws[q, bad]!r})")
        assert np.array_equal(gs[q, :n].view(np.uint32), ws[q, :n].view(np.uint32)), f"{what} query {q}: scores differ"
    if check_total:
        eq = (gr == 0) & (wr == 0)
        assert np.array_equal(gt[eq], wt[eq]), f"{what} totalHits differ"
