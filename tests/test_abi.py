"""The C-ABI library loads on a CPU-only box and exports every symbol include/nrtgpu.h declares
(no compute calls without a GPU); without a device nrtgpu_init fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "nrtgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nrtgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from nrtsearch_b200 import _native
    lib = _native.gpu_lib()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"libnrtgpu.so does not export {s}"
    assert sorted(_native.NRTGPU_SYMBOLS) == syms
    assert lib.nrtgpu_version() >= 1


def test_synth_library_loads(built):
    from nrtsearch_b200 import _native
    lib = _native.synth_lib()
    for s in ("nrtsynth_corpus_begin", "nrtsynth_corpus_fill", "nrtsynth_corpus_end", "nrtsynth_queries",
              "nrtsynth_int_column", "nrtsynth_normal_f32", "nrtsynth_uniform"):
        assert hasattr(lib, s)


def test_init_without_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nrtsearch_b200 import NrtGpuError
    from nrtsearch_b200.search import GpuContext
    with pytest.raises(NrtGpuError, match="no CPU fallback"):
        GpuContext(0)


def test_product_code_never_touches_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "nrtsearch_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(import|from)\s+oracle|#include\s+\"[^\"]*oracle", txt, flags=re.M):
                    bad.append(f)
    assert not bad, f"product files reference oracle/: {bad}"
