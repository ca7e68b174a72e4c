#!/usr/bin/env python
"""Posting-kernel time for several slice sizes (NRTGPU_SLICE_GRAN, read at context creation): python tools/slice_sweep.py DOCS 512,768,1280"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.build_if_needed()
import bench
import torch
from nrtsearch_b200 import index as ix
from nrtsearch_b200.search import GpuContext, GpuIndex, GpuIndexSearcher, RelevanceCollector
docs = int(sys.argv[1])
sh = ix.synth_text_shard(docs, 1_000_000)
sh.term_df = np.diff(sh.term_off).astype(np.int64)
sh.columns = [ix.synth_int_column(docs)]; sh.column_has = [None]
disj = bench.make_queries(1024, 1_000_000); conj = bench.make_conj_queries(1024, 1_000_000)
for cfg in sys.argv[2].split(","):
    os.environ["NRTGPU_SLICE_GRAN"] = cfg
    ctx = GpuContext(0); gix = GpuIndex(ctx, sh); s = GpuIndexSearcher(gix)
    out = {}
    for leg, (qs, thr) in {"top": (disj, 1000), "complete": (disj, 2**31 - 1), "conj": (conj, 1000)}.items():
        b = s.prepare(qs, RelevanceCollector(100, thr))
        for _ in range(2): b.run()
        torch.cuda.synchronize(); b.reset_timing()
        for _ in range(3): b.run()
        torch.cuda.synchronize()
        out[leg] = (round(b.stage_ms(0), 3), round(b.stage_ms(1), 3), b.stats()["work_items"])
        b.close()
    print("slice_gran", cfg, json.dumps(out), flush=True)
    gix.close(); ctx.close()
