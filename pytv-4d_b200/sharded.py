"""Nz-slab sharded forms of the drop-in calls: one process per GPU, every rank holds the contiguous
z-slab `[z_offset, z_offset + Nz_local)` of a `(Nz_global, M, N, N)` volume.

    sh = ShardedTV("hybrid", reg_time=2**-5)          # after torch.distributed.init_process_group("nccl")
    tv, G_slab = sh.tv(x_slab)                        # tv is the TV of the WHOLE volume (all-reduced)
    D_slab = sh.D(x_slab);  out_slab = sh.D_T(p_slab);  l21 = sh.l21(D_slab)

Exchange per call: `D` one image plane per neighbour, `D_T` one plane of the z-component(s), `tv` TWO image planes per
neighbour (the sub-gradient at a voxel needs the gradient norm of its z-neighbours, which need theirs), plus one
all-reduce of a double for the scalars.  With the z axis off (`reg_z_over_reg = 0` or a single plane in total) the
slabs are independent and nothing is exchanged.  M and N are never split.  The reference has no counterpart: it
materialises the whole gradient field on one device (tv_operators_GPU.py:175).
"""
import ctypes

import numpy as np
import torch

from . import _dev, _lib
from .cp import HaloExchange


class CudaSlabOps:
    """Executor calling libpytv_b200.so on the current stream."""

    def __init__(self):
        self.lib = _lib.lib()

    def D(self, pb, x, out, lo, hi):
        _lib.check(self.lib.pytvb_D(ctypes.byref(pb), _dev.ptr(x), _dev.ptr(out), _dev.ptr(lo), _dev.ptr(hi), _dev.stream_ptr()))

    def DT(self, pb, p, out, lo, hi):
        _lib.check(self.lib.pytvb_DT(ctypes.byref(pb), _dev.ptr(p), _dev.ptr(out), _dev.ptr(lo), _dev.ptr(hi), _dev.stream_ptr()))

    def l21(self, pb, D, Nd, d_sum):
        ws = _dev.reduce_workspace(pb, D.device)
        _lib.check(self.lib.pytvb_l21(ctypes.byref(pb), _dev.ptr(D), Nd, None, _dev.ptr(d_sum), _dev.ptr(ws), _dev.stream_ptr()))

    def apply_mask(self, pb, x, mask, is_plane):
        _lib.check(self.lib.pytvb_apply_mask(ctypes.byref(pb), _dev.ptr(x), _dev.ptr(mask), is_plane, _dev.stream_ptr()))

    def tv(self, pb, x, G, norms, d_tv, lo2, hi2):
        ws_r = _dev.reduce_workspace(pb, x.device)
        ws_t = torch.empty(self.lib.pytvb_tv_workspace_bytes(ctypes.byref(pb)), dtype=torch.uint8, device=x.device)
        _lib.check(self.lib.pytvb_tv(ctypes.byref(pb), _dev.ptr(x), _dev.ptr(G), _dev.ptr(norms), _dev.ptr(d_tv), _dev.ptr(lo2), _dev.ptr(hi2),
                                     _dev.ptr(ws_r), _dev.ptr(ws_t), _dev.stream_ptr()))

    def to_device(self, arr):
        return _dev.to_device(arr)[0]

    def mask_static(self, ms, Ni, Nj):
        return _dev.mask_static_to_device(ms, Ni, Nj)


class ShardedTV:
    def __init__(self, scheme, group=None, reg_z_over_reg=1.0, reg_time=0.0, mask_static=False, factor_reg_static=0, ops=None, comm=None):
        if scheme not in _dev.SCHEMES:
            raise ValueError("unknown scheme %r" % (scheme,))
        self.scheme = scheme
        self.rz, self.rt, self.fac = float(reg_z_over_reg), float(reg_time), float(factor_reg_static)
        self.mask_static = mask_static
        self.ops = ops if ops is not None else CudaSlabOps()
        self.halo = HaloExchange(group)
        self._layout = None
        self._ms = None
        self.transport = "nccl"

    # ---- slab placement: learned from the first call (every rank contributes its plane count)
    def _place(self, Nz_local, device):
        if self._layout is None or self._layout[0] != Nz_local:
            counts = torch.zeros(self.halo.world, dtype=torch.int64, device=device)
            counts[self.halo.rank] = Nz_local
            self.halo.allreduce_sum(counts)
            self._layout = (Nz_local, int(counts[: self.halo.rank].sum().item()), int(counts.sum().item()))
        return self._layout[1], self._layout[2]

    def _problem(self, t, shape4, device):
        z_offset, Nz_global = self._place(shape4[0], device)
        if self._ms is None and not isinstance(self.mask_static, bool):
            self._ms = self.ops.mask_static(self.mask_static, shape4[2], shape4[3])
        pb = _lib.make_problem(self.scheme, _lib.F32 if t.dtype == torch.float32 else _lib.F64, shape4, self.rz, self.rt, self.fac,
                               self._ms.data_ptr() if self._ms is not None else None, z_offset, Nz_global)
        z_on = Nz_global > 1 and self.rz > 0
        t_on = shape4[1] > 1 and self.rt > 0
        Nd = (4 + 2 * z_on + 2 * t_on) if self.scheme == "hybrid" else (2 + z_on + t_on)
        return pb, z_on, Nd

    def _planes(self, like, n):
        """(lo, hi) receive buffers of n planes each on the sides where a neighbour exists."""
        shape = ((n,) if n > 1 else ()) + tuple(like.shape[1:])
        lo = torch.empty(shape, dtype=like.dtype, device=like.device) if self.halo.prev is not None else None
        hi = torch.empty(shape, dtype=like.dtype, device=like.device) if self.halo.next is not None else None
        return lo, hi

    # ---- operators
    def D(self, x_slab):
        """D_<scheme> of this rank's slab of the whole volume: (Nz_local, M, N, N) -> (Nz_local, Nd, M, N, N)."""
        x = self.ops.to_device(x_slab)
        shape = _dev.image_shape(x)
        pb, z_on, Nd = self._problem(x, shape, x.device)
        lo = hi = None
        if z_on:
            lo, hi = self._planes(x, 1)
            self.halo.exchange(x[0], x[-1], lo, hi)
        out = torch.empty((shape[0], Nd) + shape[1:], dtype=x.dtype, device=x.device)
        self.ops.D(pb, x, out, lo, hi)
        return out

    def D_T(self, p_slab):
        """D_T_<scheme> of this rank's slab of a field: (Nz_local, Nd, M, N, N) -> (Nz_local, M, N, N)."""
        p = self.ops.to_device(p_slab)
        if p.ndim != 5:
            raise IndexError("D_T expects a 5-D field slab")
        shape = (p.shape[0],) + tuple(p.shape[2:])
        pb, z_on, Nd = self._problem(p, tuple(int(s) for s in shape), p.device)
        if Nd != p.shape[1]:
            raise IndexError("field has %d components, expected %d" % (p.shape[1], Nd))
        lo = hi = None
        if z_on:
            zf, zb = (4, 5) if self.scheme == "hybrid" else (2, 2)
            lo, hi = self._planes(p[:, 0], 1)
            # my backward-type z slot at my first plane is the previous rank's halo_hi; my forward-type slot at my last
            # plane is the next rank's halo_lo
            self.halo.exchange(p[0, zb], p[-1, zf], lo, hi)
        out = torch.empty(shape, dtype=p.dtype, device=p.device)
        self.ops.DT(pb, p, out, lo, hi)
        return out

    def l21(self, D_slab):
        """compute_L21_norm over the whole volume (all-reduced)."""
        d = self.ops.to_device(D_slab)
        shape = (d.shape[0],) + tuple(d.shape[2:])
        pb, _, _ = self._problem(d, tuple(int(s) for s in shape), d.device)
        s = torch.zeros(1, dtype=torch.float64, device=d.device)
        self.ops.l21(pb, d, int(d.shape[1]), s)
        self.halo.allreduce_sum(s)
        return float(s[0])

    def tv(self, x_slab, mask=None, return_grad_norms=False):
        """tv_<scheme>: (TV of the whole volume, sub-gradient of this slab[, gradient norms of this slab]).
        `mask` (this slab's part, same shape as the slab or one (N, N) plane) zeroes the slab in place first."""
        x = self.ops.to_device(x_slab)
        shape = _dev.image_shape(x)
        pb, z_on, Nd = self._problem(x, shape, x.device)
        if mask is not None:
            m = mask if isinstance(mask, torch.Tensor) else torch.as_tensor(np.asarray(mask))
            m = (m != 0).to(torch.uint8).to(x.device).contiguous()
            self.ops.apply_mask(pb, x, m, 1 if m.numel() == shape[2] * shape[3] else 0)
        lo2 = hi2 = None
        if z_on:
            if shape[0] < 2 and self.halo.world > 1:
                raise ValueError("sharded tv needs at least 2 planes per rank (the halo is 2 planes deep)")
            lo2, hi2 = self._planes(x, 2)
            self.halo.exchange(x[0:2], x[-2:], lo2, hi2)
        G = torch.empty_like(x)
        norms = torch.empty_like(x) if return_grad_norms else None
        s = torch.zeros(1, dtype=torch.float64, device=x.device)
        self.ops.tv(pb, x, G, norms, s, lo2, hi2)
        self.halo.allreduce_sum(s)
        return (float(s[0]), G, norms) if return_grad_norms else (float(s[0]), G)
