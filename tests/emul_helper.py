"""Loader for the host emulation of the CUDA per-quad code (tests/emul/emul.cu).  Test infrastructure."""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_DIR = os.path.join(ROOT, "tests", "emul")
EMUL_SO = os.path.join(EMUL_DIR, "libpytvb_emul.so")
sys.path.insert(0, ROOT)

import pytv_b200  # noqa: E402
from pytv_b200 import _lib  # noqa: E402


def build_emul(force=False):
    src = os.path.join(EMUL_DIR, "emul.cu")
    deps = [src] + [os.path.join(ROOT, "pytv-4d_b200", "csrc", f) for f in ("core.cuh", "strip_core.cuh", "tile_core.cuh", "tile2_core.cuh", "kernels.cuh", "kernels2.cuh", "kernels_tile.cuh", "host_common.cuh")] + [os.path.join(EMUL_DIR, "gen1_quad.cuh")]
    if not force and os.path.exists(EMUL_SO) and all(os.path.getmtime(EMUL_SO) >= os.path.getmtime(d) for d in deps):
        return EMUL_SO
    cmd = ["nvcc", "-O1", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "--extended-lambda", "-gencode",
           "arch=compute_100a,code=sm_100a", src, "-o", EMUL_SO]
    subprocess.run(cmd, check=True, cwd=EMUL_DIR)
    return EMUL_SO


GEN = 1   # 1: tests/emul/gen1_quad.cuh (retired generation-1 code), 2: strip_core.cuh, 3: strip_core.cuh with tv through the single-sweep tile kernel (tile_core.cuh) = what the library runs
_h = None


def emul():
    global _h
    if _h is None:
        _h = ctypes.CDLL(build_emul())
        VP = ctypes.c_void_p
        _h.pytvb_emulate.restype = ctypes.c_int
        _h.pytvb_emulate.argtypes = [ctypes.c_int, ctypes.POINTER(_lib.Problem), VP, VP, VP, VP, VP, VP, VP, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    return _h


def set_tv_rows(r):
    """Rows per thread of the emulated TV sweeps: 8 (what libpytv_b200.so runs) or 4."""
    emul().pytvb_emulate_set_rows(int(r))


def set_tile(strips=0, Lz=0, form=0):
    """Overrides for the emulated tile kernel: strips per frame (smaller tiles -> more CTAs along i), z-chunk length and form
    (1 = two phases per plane, tile_core.cuh; 2 = one phase, tile2_core.cuh); 0 = what the library chooses."""
    emul().pytvb_emulate_set_tile(int(strips), int(Lz), int(form))


def tile_geometry(shape, dtype=np.float32, scheme="hybrid", reg_z_over_reg=1.0, reg_time=0.0, mask_static=None, fac=0.0):
    """The tile-kernel form and geometry the library picks for a whole volume of this shape (host logic of kernels_tile.cuh)."""
    x = np.zeros((1, 1, 1, 4), dtype=dtype)       # only the dtype matters
    pb, keep = _problem(scheme, x.dtype, shape, reg_z_over_reg, reg_time, mask_static, fac)
    out = (ctypes.c_longlong * 12)()
    emul().pytvb_emulate_tile_geom(ctypes.byref(pb), out)
    keys = ("form", "strips", "RPF", "TI", "TJ", "FC", "nthreads", "Lz", "nzc", "nblocks", "smem", "smem_limit")
    return dict(zip(keys, [int(v) for v in out]))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _problem(scheme, x_dtype, shape, rz, rt, mask_static, fac, z_offset=0, Nz_global=None, time_weight=None):
    ms = None
    if not isinstance(mask_static, bool) and mask_static is not None:
        ms = np.ascontiguousarray(np.asarray(mask_static).reshape(shape[-2], shape[-1]).astype(np.uint8))
    ts = None
    if time_weight is not None:
        ts = np.ascontiguousarray(np.sqrt(np.broadcast_to(np.asarray(time_weight, dtype=np.float64), shape)).astype(x_dtype))
    pb = _lib.make_problem(scheme, _lib.F32 if x_dtype == np.float32 else _lib.F64, shape, rz, rt, fac,
                           ms.ctypes.data if ms is not None else None, z_offset, Nz_global, ts.ctypes.data if ts is not None else None)
    return pb, (ms, ts)


def _call(op, pb, inp, out, out2=None, aux=None, x0=None, lo=None, hi=None, c0=0.0, c1=0.0, variant=0, scalar=False):
    s = ctypes.c_double(0.0)
    rc = emul().pytvb_emulate(op, ctypes.byref(pb), _ptr(inp), _ptr(out), _ptr(out2), _ptr(aux), _ptr(x0), _ptr(lo), _ptr(hi), c0, c1, variant,
                              int(scalar), ctypes.byref(s))
    assert rc == 0, emul().pytvb_emulate_error()
    return s.value


def nd_of(pb):
    z_on = pb.Nz_global > 1 and pb.reg_z_over_reg > 0
    t_on = pb.M > 1 and pb.reg_time > 0
    return (4 + 2 * z_on + 2 * t_on) if pb.scheme == 3 else (2 + z_on + t_on)


def D(x, scheme, reg_z_over_reg=1.0, reg_time=0.0, mask_static=False, factor_reg_static=0.0, lo=None, hi=None, z_offset=0, Nz_global=None,
      scalar=False, time_weight=None):
    x = np.ascontiguousarray(x)
    pb, keep = _problem(scheme, x.dtype, x.shape, reg_z_over_reg, reg_time, mask_static, factor_reg_static, z_offset, Nz_global, time_weight)
    out = np.full((x.shape[0], nd_of(pb)) + x.shape[1:], np.nan, dtype=x.dtype)
    _call(0 if GEN == 1 else 7, pb, x, out, lo=lo, hi=hi, scalar=scalar)
    return out


def D_T(p, scheme, reg_z_over_reg=1.0, reg_time=0.0, mask_static=False, factor_reg_static=0.0, lo=None, hi=None, z_offset=0, Nz_global=None,
        scalar=False, time_weight=None):
    p = np.ascontiguousarray(p)
    shape = (p.shape[0],) + p.shape[2:]
    pb, keep = _problem(scheme, p.dtype, shape, reg_z_over_reg, reg_time, mask_static, factor_reg_static, z_offset, Nz_global, time_weight)
    assert nd_of(pb) == p.shape[1]
    out = np.full(shape, np.nan, dtype=p.dtype)
    _call(1 if GEN == 1 else 8, pb, p, out, lo=lo, hi=hi, scalar=scalar)
    return out


def tv(x, scheme, reg_z_over_reg=1.0, reg_time=0.0, mask_static=False, factor_reg_static=0.0, lo=None, hi=None, z_offset=0, Nz_global=None,
       scalar=False, time_weight=None, time_scale_halos=None):
    x = np.ascontiguousarray(x)
    pb, keep = _problem(scheme, x.dtype, x.shape, reg_z_over_reg, reg_time, mask_static, factor_reg_static, z_offset, Nz_global, time_weight)
    G = np.full(x.shape, np.nan, dtype=x.dtype)
    norms = np.full(x.shape, np.nan, dtype=x.dtype)
    if time_scale_halos is not None:      # (plane z = -1, plane z = Nz) of the sqrt weight map, for slabs (tile kernel)
        tlo, thi = (None if h is None else np.ascontiguousarray(np.sqrt(np.asarray(h, dtype=np.float64)).astype(x.dtype)) for h in time_scale_halos)
        pb.time_scale_lo = tlo.ctypes.data if tlo is not None else None
        pb.time_scale_hi = thi.ctypes.data if thi is not None else None
    val = _call({1: 2, 2: 9, 3: 10}[GEN], pb, x, G, out2=norms, lo=lo, hi=hi, scalar=scalar)
    return val, G, norms


def cp_dual(xbar, y, scheme, lam, sigma, lo=None, hi=None, z_offset=0, Nz_global=None, scalar=False, **w):
    pb, keep = _problem(scheme, xbar.dtype, xbar.shape, w.get("reg_z_over_reg", 1.0), w.get("reg_time", 0.0), w.get("mask_static", False),
                        w.get("factor_reg_static", 0.0), z_offset, Nz_global, w.get("time_weight"))
    return _call(3 if GEN == 1 else 5, pb, np.ascontiguousarray(xbar), y, lo=lo, hi=hi, c0=sigma, c1=1.0 / lam, scalar=scalar)


def cp_primal(y, x, aux, x0, scheme, tau, c2, variant, lo=None, hi=None, z_offset=0, Nz_global=None, scalar=False, **w):
    pb, keep = _problem(scheme, x.dtype, x.shape, w.get("reg_z_over_reg", 1.0), w.get("reg_time", 0.0), w.get("mask_static", False),
                        w.get("factor_reg_static", 0.0), z_offset, Nz_global, w.get("time_weight"))
    return _call(4 if GEN == 1 else 6, pb, np.ascontiguousarray(y), x, aux=aux, x0=x0, lo=lo, hi=hi, c0=tau, c1=c2, variant=variant, scalar=scalar)


def _mirror_protos():
    h = emul()
    VP = ctypes.c_void_p
    h.pytvb_emulate_mirror.restype = ctypes.c_int
    h.pytvb_emulate_mirror.argtypes = [ctypes.c_int, ctypes.POINTER(_lib.Problem)] + [VP] * 8 + [ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                                                                              ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    return h


def _call_mirror(op, pb, inp, out, aux, x0, lo, hi, mp, mn, c0, c1, variant, scalar):
    h = _mirror_protos()
    s = ctypes.c_double(0.0)
    rc = h.pytvb_emulate_mirror(op, ctypes.byref(pb), _ptr(inp), _ptr(out), _ptr(aux), _ptr(x0), _ptr(lo), _ptr(hi), _ptr(mp), _ptr(mn), c0, c1,
                                variant, int(scalar), ctypes.byref(s))
    assert rc == 0, h.pytvb_emulate_error()
    return s.value


def cp_dual_mirror(xbar, y, scheme, lam, sigma, lo, hi, mirror_prev, mirror_next, z_offset, Nz_global, scalar=False, **w):
    """Dual pass of one slab that also stores its boundary z-components into the neighbours' halo planes."""
    pb, keep = _problem(scheme, xbar.dtype, xbar.shape, w.get("reg_z_over_reg", 1.0), w.get("reg_time", 0.0), w.get("mask_static", False),
                        w.get("factor_reg_static", 0.0), z_offset, Nz_global)
    return _call_mirror(0, pb, np.ascontiguousarray(xbar), y, None, None, lo, hi, mirror_prev, mirror_next, sigma, 1.0 / lam, 0, scalar)


def cp_primal_mirror(y, x, aux, x0, scheme, tau, c2, variant, lo, hi, mirror_prev, mirror_next, z_offset, Nz_global, scalar=False, **w):
    pb, keep = _problem(scheme, x.dtype, x.shape, w.get("reg_z_over_reg", 1.0), w.get("reg_time", 0.0), w.get("mask_static", False),
                        w.get("factor_reg_static", 0.0), z_offset, Nz_global)
    return _call_mirror(1, pb, np.ascontiguousarray(y), x, aux, x0, lo, hi, mirror_prev, mirror_next, tau, c2, variant, scalar)


def cp_step_f16y(xbar, y_half, x, x0, scheme, lam, sigma, tau, theta, scalar=False, **w):
    """One ROF iteration with the dual field stored as normalised half (numpy float16 array); returns (l21, fid)."""
    pb, keep = _problem(scheme, x.dtype, x.shape, w.get("reg_z_over_reg", 1.0), w.get("reg_time", 0.0), w.get("mask_static", False),
                        w.get("factor_reg_static", 0.0))
    h = emul()
    VP = ctypes.c_void_p
    h.pytvb_emulate_f16y.restype = ctypes.c_int
    h.pytvb_emulate_f16y.argtypes = [ctypes.c_int, ctypes.POINTER(_lib.Problem), VP, VP, VP, VP, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_double)]
    s = ctypes.c_double(0)
    assert h.pytvb_emulate_f16y(0, ctypes.byref(pb), _ptr(xbar), _ptr(y_half), None, None, sigma, 1.0 / lam, lam, int(scalar), ctypes.byref(s)) == 0
    l21 = s.value
    assert h.pytvb_emulate_f16y(1, ctypes.byref(pb), _ptr(y_half), _ptr(x), _ptr(xbar), _ptr(x0), tau, theta, lam, int(scalar), ctypes.byref(s)) == 0
    return l21, s.value


class EmulOps:
    gen = 2

    """Executor for pytv_b200.cp.CPSolver that runs the per-quad CUDA code on the host (CPU tensors).
    Lets the multi-rank slab / halo / all-reduce logic be tested with the gloo backend."""

    @staticmethod
    def _p(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def cp_dual(self, pb, xbar, y, lam, sigma, d_l21, lo, hi, ws):
        s = ctypes.c_double(0.0)
        rc = emul().pytvb_emulate(3 if self.gen == 1 else 5, ctypes.byref(pb), self._p(xbar), self._p(y), None, None, None, self._p(lo), self._p(hi), sigma, 1.0 / lam, 0, 0,
                                  ctypes.byref(s))
        assert rc == 0
        if d_l21 is not None:
            d_l21[0] = s.value

    def cp_primal(self, variant, pb, y, x, aux, x0, tau, c2, d_fid, lo, hi, ws):
        s = ctypes.c_double(0.0)
        rc = emul().pytvb_emulate(4 if self.gen == 1 else 6, ctypes.byref(pb), self._p(y), self._p(x), None, self._p(aux), self._p(x0), self._p(lo), self._p(hi), tau, c2,
                                  0 if variant == "rof" else 1, 0, ctypes.byref(s))
        assert rc == 0
        if d_fid is not None:
            d_fid[0] = s.value

    def workspace(self, pb, device):
        import torch
        return torch.empty(1, dtype=torch.uint8)

    # the passes with mirror stores (peer-memory halo push); mirror_prev / mirror_next are raw addresses or None
    def cp_dual_p2p(self, pb, xbar, y, lam, sigma, d_l21, lo, hi, mirror_prev, mirror_next, ws):
        s = ctypes.c_double(0.0)
        _mirror_protos()
        rc = emul().pytvb_emulate_mirror(0, ctypes.byref(pb), self._p(xbar), self._p(y), None, None, self._p(lo), self._p(hi), mirror_prev, mirror_next,
                                         sigma, 1.0 / lam, 0, 0, ctypes.byref(s))
        assert rc == 0
        if d_l21 is not None:
            d_l21[0] = s.value

    def cp_primal_p2p(self, variant, pb, y, x, aux, x0, tau, c2, d_fid, lo, hi, mirror_prev, mirror_next, ws):
        s = ctypes.c_double(0.0)
        _mirror_protos()
        rc = emul().pytvb_emulate_mirror(1, ctypes.byref(pb), self._p(y), self._p(x), self._p(aux), self._p(x0), self._p(lo), self._p(hi), mirror_prev,
                                         mirror_next, tau, c2, 0 if variant == "rof" else 1, 0, ctypes.byref(s))
        assert rc == 0
        if d_fid is not None:
            d_fid[0] = s.value


class EmulSlabOps:
    """Executor for pytv_b200.sharded.ShardedTV on CPU tensors (host emulation of the strip kernels)."""

    @staticmethod
    def _p(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def _run(self, op, pb, inp, out, out2=None, lo=None, hi=None):
        s = ctypes.c_double(0.0)
        rc = emul().pytvb_emulate(op, ctypes.byref(pb), self._p(inp), self._p(out), self._p(out2), None, None, self._p(lo), self._p(hi), 0.0, 0.0, 0, 0,
                                  ctypes.byref(s))
        assert rc == 0
        return s.value

    def D(self, pb, x, out, lo, hi):
        self._run(7, pb, x, out, lo=lo, hi=hi)

    def DT(self, pb, p, out, lo, hi):
        self._run(8, pb, p, out, lo=lo, hi=hi)

    def tv(self, pb, x, G, norms, d_tv, lo2, hi2):
        d_tv[0] = self._run(10, pb, x, G, out2=norms, lo=lo2, hi=hi2)      # the tile kernel (or the fallback the library would pick)

    def l21(self, pb, D, Nd, d_sum):
        import torch
        d_sum[0] = float(torch.sqrt((D.double() ** 2).sum(dim=1)).sum())

    def apply_mask(self, pb, x, mask, is_plane):
        import torch
        m = mask.bool()
        x[~(m.expand_as(x) if is_plane else m)] = 0

    def to_device(self, arr):
        import torch
        t = arr if isinstance(arr, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(arr))
        return t.contiguous()

    def mask_static(self, ms, Ni, Nj):
        import torch
        m = ms if isinstance(ms, torch.Tensor) else torch.as_tensor(np.asarray(ms))
        return (m.reshape(Ni, Nj) != 0).to(torch.uint8).contiguous()
