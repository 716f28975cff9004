"""Nz-slab sharded forms of the drop-in calls: one process per GPU, every rank holds the contiguous
z-slab `[z_offset, z_offset + Nz_local)` of a `(Nz_global, M, N, N)` volume.

    sh = ShardedTV("hybrid", reg_time=2**-5)          # after torch.distributed.init_process_group("nccl")
    tv, G_slab = sh.tv(x_slab)                        # tv is the TV of the WHOLE volume (all-reduced)
    D_slab = sh.D(x_slab);  out_slab = sh.D_T(p_slab);  l21 = sh.l21(D_slab)

Exchange per call: `D` one image plane per neighbour, `D_T` one plane of the z-component(s), `tv` TWO image planes per
neighbour (the sub-gradient at a voxel needs the gradient norm of its z-neighbours, which need theirs) - and one plane of
the weight map when `time_weight` is given - plus one all-reduce of a double for the scalars.  With the z axis off
(`reg_z_over_reg = 0` or a single plane in total) the slabs are independent and nothing is exchanged.  M and N are never
split.  The reference has no counterpart: it materialises the whole gradient field on one device (tv_operators_GPU.py:175).

Transport (`comm`): "p2p" - every rank copies its boundary planes straight into its neighbours' halo buffers (symmetric
memory mapped over NVLink, `PeerPlanes`), fenced by a device-side handshake between neighbours, no NCCL call on the data
path; "nccl" - batched isend/irecv; "auto" (default, or PYTVB_COMM): p2p where every rank can set it up, else nccl.  With
either transport the exchange is hidden behind the interior of the slab: the planes that need no halo are computed first
while the boundary planes travel on a second stream, then the one or two boundary planes on each side follow.
"""
import ctypes
import os

import numpy as np
import torch

from . import _dev, _lib
from .cp import HaloExchange, neighbour_handshake


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


class PeerPlanes:
    """Halo planes of the stateless sharded calls in symmetric memory: per rank `[2 parities][lo, hi][nbytes]`.  A rank
    stores its boundary planes into the buffers of its z-neighbours (a device-to-device copy over NVLink into the peer
    mapping) and reads its own after a handshake with them.  Parities alternate from call to call, so the neighbour that is
    one call ahead (it cannot be further: the handshake) writes the other half while this rank still reads."""
    LO, HI = 0, 1

    def __init__(self, halo, nbytes, device):
        import torch.distributed._symmetric_memory as symm
        group = halo.group if halo.group is not None else halo.dist.group.WORLD
        self.nbytes = (int(nbytes) + 255) & ~255
        self.buf = symm.empty((2, 2, self.nbytes), dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, group)
        self.prev, self.next = halo.prev, halo.next
        self.timeout_ms = int(os.environ.get("PYTVB_P2P_TIMEOUT_MS", "60000"))
        self.parity = 0
        self.hdl.barrier(channel=0, timeout_ms=self.timeout_ms)

    def _view(self, rank, parity, side, like, nplanes_shape):
        n = int(np.prod(nplanes_shape))
        off = ((parity * 2 + side) * self.nbytes) // like.element_size()
        if rank is None:
            return self.buf[parity, side].view(like.dtype)[:n].view(nplanes_shape)
        return self.hdl.get_buffer(rank, tuple(nplanes_shape), like.dtype, off)

    def exchange(self, to_prev, to_next):
        """Store `to_prev` into the previous rank's HI buffer and `to_next` into the next rank's LO buffer, handshake; returns
        (lo, hi) views of this rank's buffers holding what the neighbours stored.  Stream-ordered on the current stream."""
        par = self.parity
        self.parity ^= 1
        if self.prev is not None and to_prev is not None:
            self._view(self.prev, par, self.HI, to_prev, tuple(to_prev.shape)).copy_(to_prev)
        if self.next is not None and to_next is not None:
            self._view(self.next, par, self.LO, to_next, tuple(to_next.shape)).copy_(to_next)
        neighbour_handshake(self.hdl, self.prev, self.next, self.timeout_ms)
        lo = self._view(None, par, self.LO, to_next, tuple(to_next.shape)) if (self.prev is not None and to_next is not None) else None
        hi = self._view(None, par, self.HI, to_prev, tuple(to_prev.shape)) if (self.next is not None and to_prev is not None) else None
        return lo, hi


class ShardedTV:
    def __init__(self, scheme, group=None, reg_z_over_reg=1.0, reg_time=0.0, mask_static=False, factor_reg_static=0, ops=None, comm=None,
                 overlap=True):
        if scheme not in _dev.SCHEMES:
            raise ValueError("unknown scheme %r" % (scheme,))
        self.scheme = scheme
        self.rz, self.rt, self.fac = float(reg_z_over_reg), float(reg_time), float(factor_reg_static)
        self.mask_static = mask_static
        self.ops = ops if ops is not None else CudaSlabOps()
        self.halo = HaloExchange(group)
        self._layout = None
        self._ms = None
        if comm is None:
            comm = os.environ.get("PYTVB_COMM", "auto")
        if comm not in ("auto", "nccl", "p2p"):
            raise ValueError("comm must be 'auto', 'nccl' or 'p2p'")
        self.comm = comm
        self.overlap = bool(overlap)
        self._peer = None            # PeerPlanes, made on first use (needs the plane size)
        self._peer_failed = False
        self._side = None            # second stream of the overlapped exchange
        self.transport = "nccl"

    # ---- slab placement: learned from the first call (every rank contributes its plane count)
    def _place(self, Nz_local, device):
        if self._layout is None or self._layout[0] != Nz_local:
            counts = torch.zeros(self.halo.world, dtype=torch.int64, device=device)
            counts[self.halo.rank] = Nz_local
            self.halo.allreduce_sum(counts)
            self._layout = (Nz_local, int(counts[: self.halo.rank].sum().item()), int(counts.sum().item()))
        return self._layout[1], self._layout[2]

    def _problem(self, t, shape4, device, ts=None, ts_lo=None, ts_hi=None, a=0, b=None):
        """Descriptor of local planes [a, b) of this rank's slab."""
        z_offset, Nz_global = self._place(shape4[0], device)
        if self._ms is None and not isinstance(self.mask_static, bool):
            self._ms = self.ops.mask_static(self.mask_static, shape4[2], shape4[3])
        b = shape4[0] if b is None else b
        sub = (b - a,) + tuple(shape4[1:])
        pb = _lib.make_problem(self.scheme, _lib.F32 if t.dtype == torch.float32 else _lib.F64, sub, self.rz, self.rt, self.fac,
                               self._ms.data_ptr() if self._ms is not None else None, z_offset + a, Nz_global,
                               ts[a].data_ptr() if ts is not None else None)
        if ts_lo is not None:
            pb.time_scale_lo = ts_lo.data_ptr()
        if ts_hi is not None:
            pb.time_scale_hi = ts_hi.data_ptr()
        z_on = Nz_global > 1 and self.rz > 0
        t_on = shape4[1] > 1 and self.rt > 0
        Nd = (4 + 2 * z_on + 2 * t_on) if self.scheme == "hybrid" else (2 + z_on + t_on)
        return pb, z_on, Nd

    # ---- halo transport
    def _peer_planes(self, nbytes, device):
        """The symmetric-memory buffers (made once, for the largest exchange seen), or None -> NCCL send/recv."""
        if self.comm == "nccl" or self._peer_failed or not isinstance(self.ops, CudaSlabOps) or self.halo.world < 2:
            return None
        if str(self.halo.dist.get_backend(self.halo.group)).lower() != "nccl":
            return None
        if self._peer is not None and self._peer.nbytes >= nbytes:
            return self._peer
        peer = None
        try:
            peer = PeerPlanes(self.halo, nbytes, device)
        except Exception as exc:
            if self.comm == "p2p":
                raise
            self._peer_error = repr(exc)
        ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=device)      # every rank must take the same path
        self.halo.dist.all_reduce(ok, op=self.halo.dist.ReduceOp.MIN, group=self.halo.group)
        if int(ok.item()) == 0:
            self._peer_failed = True
            return None
        self._peer = peer
        self.transport = "p2p"
        return peer

    def _exchange(self, to_prev, to_next):
        """Boundary planes -> neighbours; returns (lo, hi) = what the previous / next rank sent (None at the volume's ends).
        `to_prev` / `to_next`: contiguous tensors of equal shape (one or two planes, or the planes of several arrays
        concatenated)."""
        nbytes = to_prev.numel() * to_prev.element_size()
        peer = self._peer_planes(nbytes, to_prev.device)
        if peer is not None:
            return peer.exchange(to_prev, to_next)
        lo = torch.empty_like(to_next) if self.halo.prev is not None else None
        hi = torch.empty_like(to_prev) if self.halo.next is not None else None
        self.halo.exchange(to_prev, to_next, lo, hi)
        return lo, hi

    def _exchange_async(self, to_prev, to_next):
        """The exchange on a second stream (ordered after what produced the planes); returns a function that makes the
        current stream wait for it and hands out (lo, hi)."""
        if not self.overlap or not isinstance(self.ops, CudaSlabOps):
            res = self._exchange(to_prev, to_next)
            return lambda: res
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=to_prev.device)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            res = self._exchange(to_prev, to_next)
            done = torch.cuda.Event()
            done.record()

        def finish():
            cur.wait_event(done)
            for t in res:
                if t is not None:
                    t.record_stream(cur)
            return res
        return finish

    # ---- operators
    def D(self, x_slab, time_weight=None):
        """D_<scheme> of this rank's slab of the whole volume: (Nz_local, M, N, N) -> (Nz_local, Nd, M, N, N)."""
        x = self.ops.to_device(x_slab)
        shape = _dev.image_shape(x)
        ts = _dev.time_scale_to_device(time_weight, shape, x) if time_weight is not None else None
        pb, z_on, Nd = self._problem(x, shape, x.device, ts)
        out = torch.empty((shape[0], Nd) + shape[1:], dtype=x.dtype, device=x.device)
        if not z_on or self.halo.world < 2:
            self.ops.D(pb, x, out, None, None)
            return out
        Nz = shape[0]
        finish = self._exchange_async(x[0:1].contiguous(), x[Nz - 1:Nz].contiguous())
        if Nz > 2:      # interior planes need no halo: they run while the boundary planes travel
            self.ops.D(self._problem(x, shape, x.device, ts, a=1, b=Nz - 1)[0], x[1:Nz - 1], out[1:Nz - 1], x[0], x[Nz - 1])
        lo, hi = finish()
        lo = lo[0] if lo is not None else None
        hi = hi[0] if hi is not None else None
        if Nz == 1:
            self.ops.D(pb, x, out, lo, hi)
        else:
            self.ops.D(self._problem(x, shape, x.device, ts, a=0, b=1)[0], x[0:1], out[0:1], lo, x[1])
            self.ops.D(self._problem(x, shape, x.device, ts, a=Nz - 1, b=Nz)[0], x[Nz - 1:Nz], out[Nz - 1:Nz], x[Nz - 2], hi)
        return out

    def D_T(self, p_slab, time_weight=None):
        """D_T_<scheme> of this rank's slab of a field: (Nz_local, Nd, M, N, N) -> (Nz_local, M, N, N)."""
        p = self.ops.to_device(p_slab)
        if p.ndim != 5:
            raise IndexError("D_T expects a 5-D field slab")
        shape = tuple(int(s) for s in ((p.shape[0],) + tuple(p.shape[2:])))
        ts = _dev.time_scale_to_device(time_weight, shape, p) if time_weight is not None else None
        pb, z_on, Nd = self._problem(p, shape, p.device, ts)
        if Nd != p.shape[1]:
            raise IndexError("field has %d components, expected %d" % (p.shape[1], Nd))
        out = torch.empty(shape, dtype=p.dtype, device=p.device)
        if not z_on or self.halo.world < 2:
            self.ops.DT(pb, p, out, None, None)
            return out
        Nz = shape[0]
        zf, zb = (4, 5) if self.scheme == "hybrid" else (2, 2)
        # my backward-type z slot at my first plane is the previous rank's halo_hi; my forward-type slot at my last plane is
        # the next rank's halo_lo
        finish = self._exchange_async(p[0:1, zb].contiguous(), p[Nz - 1:Nz, zf].contiguous())
        if Nz > 2:
            self.ops.DT(self._problem(p, shape, p.device, ts, a=1, b=Nz - 1)[0], p[1:Nz - 1], out[1:Nz - 1], p[0, zf], p[Nz - 1, zb])
        lo, hi = finish()
        lo = lo[0] if lo is not None else None
        hi = hi[0] if hi is not None else None
        if Nz == 1:
            self.ops.DT(pb, p, out, lo, hi)
        else:
            self.ops.DT(self._problem(p, shape, p.device, ts, a=0, b=1)[0], p[0:1], out[0:1], lo, p[1, zb])
            self.ops.DT(self._problem(p, shape, p.device, ts, a=Nz - 1, b=Nz)[0], p[Nz - 1:Nz], out[Nz - 1:Nz], p[Nz - 2, zf], hi)
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

    def tv(self, x_slab, mask=None, return_grad_norms=False, time_weight=None):
        """tv_<scheme>: (TV of the whole volume, sub-gradient of this slab[, gradient norms of this slab]).
        `mask` (this slab's part, same shape as the slab or one (N, N) plane) zeroes the slab in place first.
        `time_weight`: this slab's part of a (Nz, M, N, N) weight map of the time regularisation (README.md:258)."""
        x = self.ops.to_device(x_slab)
        shape = _dev.image_shape(x)
        ts = _dev.time_scale_to_device(time_weight, shape, x) if time_weight is not None else None
        pb, z_on, Nd = self._problem(x, shape, x.device, ts)
        if mask is not None:
            m = mask if isinstance(mask, torch.Tensor) else torch.as_tensor(np.asarray(mask))
            m = (m != 0).to(torch.uint8).to(x.device).contiguous()
            self.ops.apply_mask(pb, x, m, 1 if m.numel() == shape[2] * shape[3] else 0)
        G = torch.empty_like(x)
        norms = torch.empty_like(x) if return_grad_norms else None
        Nz = shape[0]
        if not z_on or self.halo.world < 2:
            s = torch.zeros(1, dtype=torch.float64, device=x.device)
            self.ops.tv(pb, x, G, norms, s, None, None)
            self.halo.allreduce_sum(s)
            return (float(s[0]), G, norms) if return_grad_norms else (float(s[0]), G)
        if Nz < 2:
            raise ValueError("sharded tv needs at least 2 planes per rank (the halo is 2 planes deep)")
        # two image planes per neighbour, and - with a weight map - the neighbouring plane of its square root, in one message
        to_prev, to_next = x[0:2], x[Nz - 2:Nz]
        if ts is not None:
            to_prev = torch.cat([to_prev, ts[0:1]]).contiguous()
            to_next = torch.cat([to_next, ts[Nz - 1:Nz]]).contiguous()
        finish = self._exchange_async(to_prev.contiguous(), to_next.contiguous())
        s = torch.zeros(3, dtype=torch.float64, device=x.device)

        def part(a, b, lo2, hi2, slot, ts_lo=None, ts_hi=None):
            sub = self._problem(x, shape, x.device, ts, ts_lo, ts_hi, a=a, b=b)[0]
            self.ops.tv(sub, x[a:b], G[a:b], norms[a:b] if norms is not None else None, s[slot:slot + 1], lo2, hi2)

        interior = Nz >= 6 and self.overlap       # planes [2, Nz-2) read only this slab
        if interior:
            part(2, Nz - 2, x[0:2], x[Nz - 2:Nz], 0, ts[1] if ts is not None else None, ts[Nz - 2] if ts is not None else None)
        lo, hi = finish()
        lo2 = lo[0:2] if lo is not None else None
        hi2 = hi[0:2] if hi is not None else None
        tlo = lo[2] if (lo is not None and ts is not None) else None
        thi = hi[2] if (hi is not None and ts is not None) else None
        if interior:
            part(0, 2, lo2, x[2:4], 1, tlo, ts[2] if ts is not None else None)
            part(Nz - 2, Nz, x[Nz - 4:Nz - 2], hi2, 2, ts[Nz - 3] if ts is not None else None, thi)
        else:
            part(0, Nz, lo2, hi2, 0, tlo, thi)
        tot = s.sum().reshape(1)
        self.halo.allreduce_sum(tot)
        return (float(tot[0]), G, norms) if return_grad_norms else (float(tot[0]), G)
