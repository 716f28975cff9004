"""Drop-in for `pytv.tv_GPU` (reference pytv/tv_GPU.py): tv_hybrid / tv_downwind / tv_upwind / tv_central
return (tv, subgradient[, grad_norms]) with the reference's conventions:
  * `tv` is a 0-d numpy array (tv_operators_GPU.py:87 via tv_GPU.py:85);
  * `G` (and `grad_norms`) are numpy arrays unless return_pytorch_tensor=True, also for tensor input
    (tv_GPU.py:129-139);
  * `mask` zeroes the caller's image IN PLACE outside the mask (tv_GPU.py:79-80); for a CUDA tensor that is a
    device kernel on the caller's storage, for a numpy array the masked image is written back into it
    (set MASK_WRITEBACK = False to skip that copy);
  * zero gradient norms give a zero sub-gradient contribution and `inf` in grad_norms (tv_GPU.py:88).
"""
import ctypes

import numpy as np
import torch

from . import _dev, _lib

MASK_WRITEBACK = True
PLAN_CACHE_MAX_VOXELS = 1 << 22      # numpy-in / numpy-out calls on volumes up to this size reuse a cached TVPlan (launch-bound regime)
PLAN_CACHE_SIZE = 8
_plan_cache = {}


class TVPlan:
    """tv_<scheme> prepared for repeated calls on one shape (the README's descent loop calls it 300 times, README.md:120): the
    input buffer `x`, the outputs `G` (and `norms`), the scalar and the scratch are allocated once and the launch is captured
    in a CUDA graph, so a call is one graph replay - no allocation, no host round trip until the value is read.

        plan = TVPlan("hybrid", (1, 1, 256, 256), torch.float64)
        plan.x.copy_(img);  tv_dev, G = plan.run()        # tv_dev: 0-d float64 device tensor; G is plan-owned (overwritten by the next run)
        tv, G = plan(img)                                # copy in + run

    The sub-gradient and the value are those of tv_<scheme> (tv_GPU.py:47,142,217,290)."""

    def __init__(self, scheme, shape, dtype=torch.float32, device=None, reg_z_over_reg=1.0, reg_time=0.0, mask_static=False, factor_reg_static=0,
                 time_weight=None, return_grad_norms=False, graph=True):
        _dev.require_cuda()
        if scheme not in _dev.SCHEMES:
            raise ValueError("unknown scheme %r" % (scheme,))
        self.lib = _lib.lib()
        self.shape = tuple(int(s) for s in shape)
        if len(self.shape) != 4:
            raise IndexError("TVPlan expects a 4-D shape (Nz, M, N, N)")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        dtype = torch.float32 if dtype in (torch.float32, np.float32) else torch.float64
        self.x = torch.zeros(self.shape, dtype=dtype, device=device)
        self.G = torch.empty_like(self.x)
        self.norms = torch.empty_like(self.x) if return_grad_norms else None
        self.d_tv = torch.zeros(1, dtype=torch.float64, device=device)
        self._ms = _dev.mask_static_to_device(mask_static, self.shape[2], self.shape[3])
        self._ts = _dev.time_scale_to_device(time_weight, self.shape, self.x)
        self.pb = _dev.problem(scheme, self.x, self.shape, reg_z_over_reg, reg_time, factor_reg_static, self._ms, ts=self._ts)
        self._ws_r = _dev.reduce_workspace(self.pb, device)
        self._ws_t = torch.empty(self.lib.pytvb_tv_workspace_bytes(ctypes.byref(self.pb)), dtype=torch.uint8, device=device)
        self._use_graph = bool(graph)
        self._graph = None

    def _launch(self):
        _lib.check(self.lib.pytvb_tv(ctypes.byref(self.pb), _dev.ptr(self.x), _dev.ptr(self.G), _dev.ptr(self.norms), _dev.ptr(self.d_tv), None, None,
                                     _dev.ptr(self._ws_r), _dev.ptr(self._ws_t), _dev.stream_ptr()))

    def run(self):
        """tv and sub-gradient of the current contents of `self.x`; returns (0-d device tensor, G)."""
        if not self._use_graph or torch.cuda.is_current_stream_capturing():
            self._launch()
        else:
            if self._graph is None:
                side = torch.cuda.Stream(device=self.x.device)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._launch()          # warm-up outside the capture (lazy module loading, function attributes)
                torch.cuda.current_stream().wait_stream(side)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch()
                self._graph = g
            self._graph.replay()
        return self.d_tv[0], self.G

    def __call__(self, img):
        src = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(img))
        self.x.copy_(src, non_blocking=True)
        return self.run()


def _cached_plan(scheme, shape, dtype, device, reg_z_over_reg, reg_time, factor_reg_static, return_grad_norms):
    key = (scheme, tuple(shape), dtype, str(device), float(reg_z_over_reg), float(reg_time), float(factor_reg_static), bool(return_grad_norms))
    plan = _plan_cache.pop(key, None)
    if plan is None:
        plan = TVPlan(scheme, shape, dtype, device, reg_z_over_reg, reg_time, False, factor_reg_static, None, return_grad_norms)
        while len(_plan_cache) >= PLAN_CACHE_SIZE:
            _plan_cache.pop(next(iter(_plan_cache)))
    _plan_cache[key] = plan          # most recently used last
    return plan


def _has_mask(mask):
    return not (isinstance(mask, list) and len(mask) == 0) and mask is not None


def _mask_to_device(mask, shape):
    m = mask if isinstance(mask, torch.Tensor) else torch.as_tensor(np.asarray(mask))
    m = (m != 0)
    if tuple(m.shape) == tuple(shape):
        return m.to(torch.uint8).cuda().contiguous(), 0
    if m.numel() == shape[2] * shape[3] and tuple(m.shape[-2:]) == tuple(shape[2:]):
        return m.reshape(shape[2], shape[3]).to(torch.uint8).cuda().contiguous(), 1
    return torch.broadcast_to(m, shape).to(torch.uint8).cuda().contiguous(), 0


def _tv(scheme, img, mask, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, return_grad_norms, time_weight=None):
    shape = _dev.image_shape(img)
    # the reference's default call - numpy in, numpy out, no mask - on a launch-bound volume: a cached plan (pre-allocated
    # buffers, one graph replay) instead of five allocations and four launches per call
    if (isinstance(img, np.ndarray) and not return_pytorch_tensor and not _has_mask(mask) and isinstance(mask_static, bool)
            and time_weight is None and int(np.prod(shape)) <= PLAN_CACHE_MAX_VOXELS and PLAN_CACHE_SIZE > 0):
        _dev.require_cuda()
        dt = torch.float32 if img.dtype == np.float32 else torch.float64
        plan = _cached_plan(scheme, shape, dt, torch.device("cuda", torch.cuda.current_device()), reg_z_over_reg, reg_time, factor_reg_static,
                            return_grad_norms)
        src = torch.from_numpy(np.ascontiguousarray(img))
        plan.x.copy_(src if src.dtype == dt else src.to(dt), non_blocking=True)
        d_tv, G = plan.run()
        tv = d_tv.to(dt).cpu().numpy()
        if return_grad_norms:
            return (tv, G.cpu().numpy(), plan.norms.cpu().numpy())
        return (tv, G.cpu().numpy())
    x, was_tensor = _dev.to_device(img)
    lib = _lib.lib()
    ms = _dev.mask_static_to_device(mask_static, shape[2], shape[3])
    ts = _dev.time_scale_to_device(time_weight, shape, x)
    pb = _dev.problem(scheme, x, shape, reg_z_over_reg, reg_time, factor_reg_static, ms, ts=ts)
    st = _dev.stream_ptr()
    if _has_mask(mask):
        m, is_plane = _mask_to_device(mask, shape)
        _lib.check(lib.pytvb_apply_mask(ctypes.byref(pb), _dev.ptr(x), _dev.ptr(m), is_plane, st))
        # honour the reference's side effect on the caller's array
        shares_storage = was_tensor and img.is_cuda and img.dtype == x.dtype and img.is_contiguous()
        if MASK_WRITEBACK and not shares_storage:
            if was_tensor:
                img.copy_(x.to(img.dtype))
            else:
                img[...] = x.cpu().numpy().astype(img.dtype, copy=False)
    G = torch.empty(shape, dtype=x.dtype, device=x.device)
    norms = torch.empty(shape, dtype=x.dtype, device=x.device) if return_grad_norms else None
    d_tv = torch.empty(1, dtype=torch.float64, device=x.device)
    ws_r = _dev.reduce_workspace(pb, x.device)
    ws_t = torch.empty(lib.pytvb_tv_workspace_bytes(ctypes.byref(pb)), dtype=torch.uint8, device=x.device)
    _lib.check(lib.pytvb_tv(ctypes.byref(pb), _dev.ptr(x), _dev.ptr(G), _dev.ptr(norms), _dev.ptr(d_tv), None, None, _dev.ptr(ws_r),
                            _dev.ptr(ws_t), st))
    tv = d_tv[0].to(x.dtype).cpu().numpy()
    if return_grad_norms:
        return (tv, _dev.to_output(G, return_pytorch_tensor), _dev.to_output(norms, return_pytorch_tensor))
    return (tv, _dev.to_output(G, return_pytorch_tensor))


def tv_hybrid(img, mask=[], reg_z_over_reg=1.0, reg_time=0.0, mask_static=False, factor_reg_static=0, return_pytorch_tensor=False,
              return_grad_norms=False, time_weight=None):
    """TV value and sub-gradient, hybrid discretisation (tv_GPU.py:47)."""
    return _tv("hybrid", img, mask, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, return_grad_norms, time_weight)


def tv_downwind(img, mask=[], reg_z_over_reg=1.0, reg_time=0.0, mask_static=False, factor_reg_static=0, return_pytorch_tensor=False,
                return_grad_norms=False, time_weight=None):
    """TV value and sub-gradient, downwind discretisation (tv_GPU.py:142)."""
    return _tv("downwind", img, mask, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, return_grad_norms, time_weight)


def tv_upwind(img, mask=[], reg_z_over_reg=1.0, reg_time=0.0, mask_static=False, factor_reg_static=0, return_pytorch_tensor=False,
              return_grad_norms=False, time_weight=None):
    """TV value and sub-gradient, upwind discretisation (tv_GPU.py:217)."""
    return _tv("upwind", img, mask, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, return_grad_norms, time_weight)


def tv_central(img, mask=[], reg_z_over_reg=1.0, reg_time=0.0, mask_static=False, factor_reg_static=0, return_pytorch_tensor=False,
               return_grad_norms=False, time_weight=None):
    """TV value and sub-gradient, central discretisation (tv_GPU.py:290)."""
    return _tv("central", img, mask, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, return_grad_norms, time_weight)
