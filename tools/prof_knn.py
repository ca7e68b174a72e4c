#!/usr/bin/env python
"""Stage timings of the kNN path on the configs[3] shape without the oracle gate (for kernel experiments / ncu captures):
python tools/prof_knn.py [--vectors 1000000] [--dims 768] [--nq 1024] [--k 100] [--runs 5]"""
import argparse, ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vectors", type=int, default=1_000_000); ap.add_argument("--dims", type=int, default=768)
    ap.add_argument("--nq", type=int, default=1024); ap.add_argument("--k", type=int, default=100); ap.add_argument("--runs", type=int, default=5)
    a = ap.parse_args()
    import __graft_entry__ as g
    g.build_if_needed()
    from nrtsearch_b200 import _native, index as ix
    from nrtsearch_b200.index import HostShard
    from nrtsearch_b200.search import GpuContext, GpuIndex
    corpus = ix.synth_vectors(a.vectors, a.dims)
    q = ix.synth_vectors(a.nq, a.dims, seed=ix.SEED_VQUERIES)
    sh = HostShard(n_docs=a.vectors, doc_base=0, term_off=np.zeros(1, np.int64), post_docs=np.zeros(0, np.int32), post_freqs=np.zeros(0, np.int32),
                   fields=[], vectors=corpus, vec_similarity=ix.SIM_COSINE)
    ctx = GpuContext(0); gix = GpuIndex(ctx, sh); lib = _native.gpu_lib()
    docs, scores, counts = np.zeros((a.nq, a.k), np.int32), np.zeros((a.nq, a.k), np.float32), np.zeros(a.nq, np.int32)
    st = (ctypes.c_float * 3)()
    out = []
    for i in range(a.runs + 2):
        _native.check(lib.nrtgpu_search_knn_timed(gix.handle, q.ctypes.data, a.nq, a.k, None, docs.ctypes.data, scores.ctypes.data, counts.ctypes.data, st))
        if i >= 2:
            out.append((st[0], st[1], st[2]))
    m = np.mean(out, axis=0)
    print(json.dumps({"gemm_ms": float(m[0]), "select_ms": float(m[1]), "rescore_ms": float(m[2]),
                      "gemm_tflops": 2.0 * a.nq * a.vectors * a.dims / (m[0] * 1e-3) / 1e12}))
    gix.close(); ctx.close()


if __name__ == "__main__":
    main()
