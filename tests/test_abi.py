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
    assert lib.dabb_abi_version() == 3


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


def test_null_context_is_rejected_everywhere():
    """error convention of the ABI: int status, never a crash - every entry point refuses a NULL context / NULL arguments"""
    pkg = load_pkg()
    lib = pkg.load_library()
    null = C.c_void_p(None)
    E_ARG = -3
    calls = {
        "dabb_stream_reset": (null, 0, 1, C.c_int64(0)), "dabb_set_options": (null, null), "dabb_get_info": (null, 0, null),
        "dabb_select_subchannel": (null, 0, 1, 0, null), "dabb_remove_subchannel": (null, 0, 1, 0),
        "dabb_process": (null, null), "dabb_process_async": (null, null), "dabb_sync": (null,), "dabb_join_lanes": (null,), "dabb_submit": (null, null), "dabb_collect": (null,),
        "dabb_profile": (null, 1), "dabb_profile_read": (null, null, C.c_size_t(0)), "dabb_read_tap": (null, 0, null, C.c_size_t(0)),
        "dabb_ofdm_demod": (null, null, C.c_int64(0), null, 1, null, null, null, null),
        "dabb_find_index": (null, null, C.c_int64(0), null, 1, null, null), "dabb_find_index_ex": (null, null, C.c_int64(0), null, 1, 0, null, null),
        "dabb_coarse_estimate": (null, null, C.c_int64(0), null, 1, 0, null),
        "dabb_viterbi": (null, null, 1, 768, null), "dabb_fic_decode": (null, null, 1, null, null),
        "dabb_msc_decode": (null, null, null, 1, null), "dabb_rs_superframes": (null, null, 1, 1440, null),
        "dabb_dev_alloc": (null, C.c_size_t(16), null), "dabb_dev_free": (null, null),
        "dabb_memcpy_h2d": (null, null, null, C.c_size_t(0)), "dabb_memcpy_d2h": (null, null, null, C.c_size_t(0)),
    }
    for name, args in calls.items():
        rc = getattr(lib, name)(*args)
        assert rc < 0, (name, rc)
    assert calls.keys() <= set(pkg.EXPORTS)
    lib.dabb_destroy(null)                       # a no-op
    assert lib.dabb_kernel_launches(null) == 0
    assert lib.dabb_cuda_stream(null) is None
    h = C.c_void_p()
    assert lib.dabb_create(None, C.byref(h)) == E_ARG and lib.dabb_create(C.byref(pkg.dabb200.Config()), None) == E_ARG
