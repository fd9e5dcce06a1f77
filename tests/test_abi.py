"""The C-ABI library loads and exports every entry point include/dab_b200.h declares; without a CUDA device it refuses to
create a context (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, load_pkg


def header_functions():
    src = open(os.path.join(ROOT, "include", "dab_b200.h")).read()
    return sorted(set(re.findall(r"\b(dabb_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header():
    pkg = load_pkg()
    if not os.path.exists(pkg.LIB_PATH):
        pkg.build()
    lib = pkg.load_library()
    fns = header_functions()
    assert len(fns) >= 20
    for f in fns:
        assert hasattr(lib, f), f"{f} declared in include/dab_b200.h but not exported"
    assert sorted(pkg.EXPORTS) == fns
    assert lib.dabb_abi_version() == 2


def test_struct_sizes():
    pkg = load_pkg()
    from welle_io_b200 import dabb200 as d
    assert C.sizeof(d.Config) == 64 and C.sizeof(d.Subchannel) == 36 and C.sizeof(d.Options) == 32
    assert C.sizeof(d.FrameResult) == d.RESULT_DTYPE.itemsize == 224


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    pkg = load_pkg()
    with pytest.raises(pkg.DabbError, match="no CUDA device|CUDA"):
        pkg.Context(n_streams=1)
