"""The C-ABI library loads on a box without a GPU, exports every symbol include/fdga.h declares, and fails loudly
(no CPU fallback) when asked to create a context without a CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    import fddgasolver_jl_b200 as fd
    lib = fd._lib.load()
    hdr = open(os.path.join(ROOT, "include", "fdga.h")).read()
    declared = set(re.findall(r"\b(fdga_[a-zA-Z0-9_]+)\s*\(", hdr))
    declared -= {"fdga_ctx"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/fdga.h but not exported by libfdga.so"
    assert declared == set(fd.EXPORTS), declared ^ set(fd.EXPORTS)


def test_create_fails_loudly_without_gpu():
    import fddgasolver_jl_b200 as fd
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(fd.FdgaError):
        fd.parquet_solver_hubbard_parquet_approximation_NL2(4, 4, (2, 2), (2, 2), 4, 2, T=0.5, U=1.0, μ=0.0, t1=1.0)


def test_product_never_imports_oracle():
    """the shipped package must not reference oracle/ (a product path through the oracle voids parity)"""
    pkg = os.path.join(ROOT, "fddgasolver.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "fdga_oracle" not in txt and "orc_" not in txt, os.path.join(dirpath, f)
