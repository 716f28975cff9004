"""The drop-in boundary: libpytv_b200.so loads, exports exactly what include/pytv_b200.h declares, and its
argument checking works without a GPU (no compute call is made here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import pytv_b200
from pytv_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pytv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pytvb_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    names = _declared()
    assert len(names) >= 18
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), "%s declared in the header but not exported" % n
    assert sorted(_lib._PROTOTYPES) == names, "ctypes prototypes and header drifted apart"
    assert _lib.lib().pytvb_version() == 101


def test_build_id_names_the_sources():
    """pytvb_build_id() = hash of the sources the library was built from (what profiles/traffic.json is keyed by): 16 hex digits, and
    equal to the hash of the sources in this tree when the library is fresh (build() asserts the same)."""
    import re
    import sys
    bid = _lib.lib().pytvb_build_id().decode()
    assert re.fullmatch(r"[0-9a-f]{16}", bid), bid
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as g
    assert bid == g.source_hash()


def test_problem_struct_layout_matches_header():
    # 2 x int32, 6 x int64, 3 x double, 4 pointers, no padding surprises
    assert ctypes.sizeof(_lib.Problem) == 8 + 6 * 8 + 3 * 8 + 4 * 8
    assert _lib.Problem.Nz.offset == 8 and _lib.Problem.reg_z_over_reg.offset == 56 and _lib.Problem.mask_static.offset == 80
    assert _lib.Problem.time_scale.offset == 88 and _lib.Problem.time_scale_lo.offset == 96 and _lib.Problem.time_scale_hi.offset == 104


@pytest.mark.parametrize("scheme,Nz,M,rz,rt,expect", [
    ("hybrid", 20, 4, 1.0, 0.0, 6), ("hybrid", 20, 4, 1.0, 2 ** -5, 8), ("hybrid", 1, 1, 1.0, 1.0, 4), ("hybrid", 1, 3, 1.0, 1.0, 6),
    ("upwind", 20, 4, 1.0, 0.0, 3), ("downwind", 20, 4, 0.0, 1.0, 3), ("central", 20, 4, 0.5, 0.5, 4), ("central", 1, 1, 1.0, 0.0, 2),
    ("upwind", 5, 2, float("nan"), 0.0, 2),
])
def test_num_components(scheme, Nz, M, rz, rt, expect):
    from oracle import tv_oracle as orc
    pb = _lib.make_problem(scheme, _lib.F32, (Nz, M, 8, 8), rz, rt)
    assert _lib.lib().pytvb_num_components(ctypes.byref(pb)) == expect
    assert orc.num_components(scheme, Nz, M, rz, rt) == expect


def test_argument_errors_are_reported():
    lib = _lib.lib()
    bad = _lib.make_problem("hybrid", _lib.F32, (0, 1, 8, 8))
    assert lib.pytvb_num_components(ctypes.byref(bad)) == -1
    assert b"empty volume" in lib.pytvb_last_error()
    pb = _lib.make_problem("hybrid", _lib.F32, (4, 1, 8, 8))
    assert lib.pytvb_D(ctypes.byref(pb), None, None, None, None, None) == -1
    assert b"NULL" in lib.pytvb_last_error()
    # a slab strictly inside the volume needs both halos for the hybrid scheme
    slab = _lib.make_problem("hybrid", _lib.F32, (2, 1, 8, 8), z_offset=1, Nz_global=4)
    buf = np.zeros(2 * 6 * 64, np.float32)
    p = buf.ctypes.data_as(ctypes.c_void_p)
    assert lib.pytvb_D(ctypes.byref(slab), p, p, None, None, None) == -1
    assert b"halo_lo is required" in lib.pytvb_last_error()
    huge = _lib.make_problem("hybrid", _lib.F32, (1, 1, 1 << 16, 1 << 16))
    assert lib.pytvb_num_components(ctypes.byref(huge)) == -1 and b"2^31" in lib.pytvb_last_error()
    outside = _lib.make_problem("hybrid", _lib.F32, (4, 1, 8, 8), z_offset=2, Nz_global=4)
    assert lib.pytvb_num_components(ctypes.byref(outside)) == -1
    with pytest.raises(_lib.PytvError):
        _lib.check(-1)


def test_workspace_sizes():
    lib = _lib.lib()
    pb = _lib.make_problem("hybrid", _lib.F32, (20, 4, 100, 100), reg_time=2 ** -5)
    r = lib.pytvb_reduce_workspace_bytes(ctypes.byref(pb))
    t = lib.pytvb_tv_workspace_bytes(ctypes.byref(pb))
    assert r >= 8 * (256 + 20 * 4 * 100 * 100 // 256)
    assert 0 < t <= 4096                      # the single-sweep kernel needs no inverse-norm field
    many = _lib.make_problem("hybrid", _lib.F32, (3, 20, 16, 16), reg_time=1.0)     # 20 coupled frames: the two-sweep fallback
    assert lib.pytvb_tv_workspace_bytes(ctypes.byref(many)) >= 5 * 20 * 16 * 16 * 4


def test_partition_z():
    assert pytv_b200.partition_z(1024, 8) == [(128 * r, 128) for r in range(8)]
    assert pytv_b200.partition_z(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]
    with pytest.raises(ValueError):
        pytv_b200.partition_z(3, 4)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        pytv_b200.tv_GPU.tv_hybrid(np.zeros((1, 1, 4, 4)))
    with pytest.raises(RuntimeError, match="no CPU path"):
        pytv_b200.tv_operators_GPU.D_upwind(np.zeros((1, 1, 4, 4)))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pytv-4d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "tv_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
