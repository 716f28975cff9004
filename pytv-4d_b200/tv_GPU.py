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
