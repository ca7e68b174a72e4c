"""N > 1 host logic on CPU: world_size-2 gloo run of the shard build (doc-range split of one synthetic corpus,
all-reduced index-wide statistics) + the per-step all-gather layout. The per-shard searches and the final
TopDocs.merge are done by the oracle here (no GPU); the result must equal the single-index search bit for bit."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N_DOCS, VOCAB, NQ, K = 40_000, 4_000, 24, 20


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import oracle
    from nrtsearch_b200 import index as ix
    from nrtsearch_b200.search import BooleanQuery, Occur, TermQuery, compile_queries
    from nrtsearch_b200.shards import TopKGather, install_global_stats, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(N_DOCS, rank, world)
    sh = ix.synth_text_shard(hi - lo, VOCAB, doc_begin=lo)
    install_global_stats(sh)
    terms = ix.synth_query_terms(NQ, 3, VOCAB, log10_lo=0.3, log10_hi=3.0)
    qs = [BooleanQuery().add(TermQuery(int(t[0])), Occur.SHOULD).add(TermQuery(int(t[1])), Occur.MUST if i % 3 == 0 else Occur.SHOULD)
          .add(TermQuery(int(t[2])), Occur.SHOULD) for i, t in enumerate(terms)]
    d, s, c, tot, _ = oracle.search_compiled(oracle.OracleIndex(sh), *compile_queries(qs), K)
    g = TopKGather(NQ, K, world, torch.device("cpu"))
    g.loc_docs.copy_(torch.from_numpy(d.reshape(-1))); g.loc_scores.copy_(torch.from_numpy(s.reshape(-1)))
    g.loc_counts.copy_(torch.from_numpy(c))
    g.gather()
    tt = torch.from_numpy(tot.copy()); dist.all_reduce(tt)
    if rank == 0:
        md, ms, mc = oracle.merge_topk(g.all_docs.numpy().reshape(world, NQ, K), g.all_scores.numpy().reshape(world, NQ, K),
                                       g.all_counts.numpy().reshape(world, NQ), K)
        np.savez(out_path, docs=md, scores=ms, counts=mc, total=tt.numpy(), df=sh.term_df, ttf=sh.fields[0].sum_total_term_freq,
                 dc=sh.fields[0].doc_count)
    dist.destroy_process_group()


def test_two_rank_sharded_search_equals_single_index(built, tmp_path):
    import oracle
    from nrtsearch_b200 import index as ix
    from nrtsearch_b200.search import BooleanQuery, Occur, TermQuery, compile_queries
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    whole = ix.synth_text_shard(N_DOCS, VOCAB)
    assert np.array_equal(got["df"], np.diff(whole.term_off)) and int(got["ttf"]) == whole.fields[0].sum_total_term_freq
    assert int(got["dc"]) == N_DOCS
    terms = ix.synth_query_terms(NQ, 3, VOCAB, log10_lo=0.3, log10_hi=3.0)
    qs = [BooleanQuery().add(TermQuery(int(t[0])), Occur.SHOULD).add(TermQuery(int(t[1])), Occur.MUST if i % 3 == 0 else Occur.SHOULD)
          .add(TermQuery(int(t[2])), Occur.SHOULD) for i, t in enumerate(terms)]
    d, s, c, tot, _ = oracle.search_compiled(oracle.OracleIndex(whole), *compile_queries(qs), K)
    assert np.array_equal(got["counts"], c) and np.array_equal(got["docs"], d)
    assert np.array_equal(got["scores"].view(np.uint32), s.view(np.uint32)) and np.array_equal(got["total"], tot)
