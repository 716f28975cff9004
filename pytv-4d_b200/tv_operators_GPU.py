"""Drop-in for `pytv.tv_operators_GPU` (reference pytv/tv_operators_GPU.py): same function names, keyword
arguments, shapes and return conventions; the work is done by the sm_100a kernels of libpytv_b200.so.

Return conventions kept from the reference:
  * numpy in -> numpy out unless return_pytorch_tensor=True; tensor in -> CUDA tensor out always
    (tv_operators_GPU.py:178-182, :246-249);
  * float32 stays float32, every other dtype is computed in float64 (`type_like`, :92-131);
  * compute_L21_norm returns a 0-d numpy array, and with return_array=True the norm array is a CUDA tensor
    even when return_pytorch_tensor=False (:83-90).
Extension: every operator also takes `time_weight=None`, a (Nz, M, N, N) weight map of the time regularisation (the
reference's TODO, README.md:258); mask_static / factor_reg_static keep working and multiply on top of it.
Deliberate differences: outputs are allocated on the device (the reference allocates zeros on the host
and uploads them, :175); float32 numpy input gives float32 output also for hybrid/central (the reference
upcasts by accident under numpy >= 2); Ni != Nj is accepted.
"""
import ctypes

import numpy as np
import torch

from . import _dev, _lib


def compute_L21_norm(D_img, return_array=False, return_pytorch_tensor=False):
    """|x|_2,1 = sum_i sqrt(sum_j x_ij^2) of a (Nz, Nd, M, N, N) field (tv_operators_GPU.py:46)."""
    if D_img.ndim != 5:
        raise IndexError("compute_L21_norm expects a 5-D field (Nz, Nd, M, N, N)")
    d, _ = _dev.to_device(D_img)
    Nz, Nd, M, Ni, Nj = (int(s) for s in d.shape)
    pb = _dev.problem("upwind", d, (Nz, M, Ni, Nj), 0.0, 0.0, 0.0, None)
    ws = _dev.reduce_workspace(pb, d.device)
    out = torch.empty(1, dtype=torch.float64, device=d.device)
    norms = torch.empty((Nz, M, Ni, Nj), dtype=d.dtype, device=d.device) if return_array else None
    _lib.check(_lib.lib().pytvb_l21(ctypes.byref(pb), _dev.ptr(d), Nd, _dev.ptr(norms), _dev.ptr(out), _dev.ptr(ws), _dev.stream_ptr()))
    l21 = out[0].to(d.dtype)
    # the reference's return quirks (tv_operators_GPU.py:83-90): without return_array the scalar is ALWAYS a
    # 0-d numpy array; with return_array it is a tensor only if return_pytorch_tensor, and the norm array is a tensor always
    if return_array:
        return (l21 if return_pytorch_tensor else l21.cpu().numpy(), norms)
    return l21.cpu().numpy()


def type_like(array, array_ref):
    """`array` converted to the dtype policy of `array_ref`: float32 if the reference array is float32, else
    float64 (tv_operators_GPU.py:92-131)."""
    ref_is_f32 = array_ref.dtype in (np.float32, torch.float32)
    if isinstance(array, np.ndarray):
        if isinstance(array_ref, np.ndarray):
            return array.astype(array_ref.dtype)
        return array.astype(np.float32 if ref_is_f32 else np.float64)
    return array.type(torch.float32 if ref_is_f32 else torch.float64)


def _forward(scheme, img, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, time_weight=None):
    shape = _dev.image_shape(img)
    x, was_tensor = _dev.to_device(img)
    ms = _dev.mask_static_to_device(mask_static, shape[2], shape[3])
    ts = _dev.time_scale_to_device(time_weight, shape, x)
    pb = _dev.problem(scheme, x, shape, reg_z_over_reg, reg_time, factor_reg_static, ms, ts=ts)
    Nd = _lib.lib().pytvb_num_components(ctypes.byref(pb))
    if Nd < 0:
        _lib.check(Nd)
    out = torch.empty((shape[0], Nd) + shape[1:], dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().pytvb_D(ctypes.byref(pb), _dev.ptr(x), _dev.ptr(out), None, None, _dev.stream_ptr()))
    return _dev.to_output(out, return_pytorch_tensor or was_tensor)


def _adjoint(scheme, field, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, time_weight=None):
    if field.ndim != 5:
        raise IndexError("D_T expects a 5-D field (Nz, Nd, M, N, N); got %d dimensions" % field.ndim)
    p, was_tensor = _dev.to_device(field)
    Nz, Nd, M, Ni, Nj = (int(s) for s in p.shape)
    ms = _dev.mask_static_to_device(mask_static, Ni, Nj)
    ts = _dev.time_scale_to_device(time_weight, (Nz, M, Ni, Nj), p)
    pb = _dev.problem(scheme, p, (Nz, M, Ni, Nj), reg_z_over_reg, reg_time, factor_reg_static, ms, ts=ts)
    expect = _lib.lib().pytvb_num_components(ctypes.byref(pb))
    if expect < 0:
        _lib.check(expect)
    if expect != Nd:
        raise IndexError("field has %d components but D_T_%s with these weights acts on %d" % (Nd, scheme, expect))
    out = torch.empty((Nz, M, Ni, Nj), dtype=p.dtype, device=p.device)
    _lib.check(_lib.lib().pytvb_DT(ctypes.byref(pb), _dev.ptr(p), _dev.ptr(out), None, None, _dev.stream_ptr()))
    return _dev.to_output(out, return_pytorch_tensor or was_tensor)


def D_hybrid(img, reg_z_over_reg=1.0, reg_time=0, mask_static=False, factor_reg_static=0, return_pytorch_tensor=False, time_weight=None):
    """(Nz,M,N,N) -> (Nz,Nd,M,N,N), hybrid scheme, Nd = 4/6/8 (tv_operators_GPU.py:134)."""
    return _forward("hybrid", img, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, time_weight)


def D_downwind(img, reg_z_over_reg=1.0, reg_time=0, mask_static=False, factor_reg_static=0, return_pytorch_tensor=False, time_weight=None):
    """Backward differences, Nd = 2/3/4 (tv_operators_GPU.py:253)."""
    return _forward("downwind", img, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, time_weight)


def D_upwind(img, reg_z_over_reg=1.0, reg_time=0, mask_static=False, factor_reg_static=0, return_pytorch_tensor=False, time_weight=None):
    """Forward differences, Nd = 2/3/4 (tv_operators_GPU.py:362)."""
    return _forward("upwind", img, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, time_weight)


def D_central(img, reg_z_over_reg=1.0, reg_time=0, mask_static=False, factor_reg_static=0, return_pytorch_tensor=False, time_weight=None):
    """Centred differences / 2, Nd = 2/3/4 (tv_operators_GPU.py:471)."""
    return _forward("central", img, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, time_weight)


def D_T_hybrid(img, reg_z_over_reg=1.0, reg_time=0, mask_static=False, factor_reg_static=0, return_pytorch_tensor=False, time_weight=None):
    """(Nz,Nd,M,N,N) -> (Nz,M,N,N), adjoint of D_hybrid (tv_operators_GPU.py:583)."""
    return _adjoint("hybrid", img, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, time_weight)


def D_T_downwind(img, reg_z_over_reg=1.0, reg_time=0, mask_static=False, factor_reg_static=0, return_pytorch_tensor=False, time_weight=None):
    """Adjoint of D_downwind (tv_operators_GPU.py:719)."""
    return _adjoint("downwind", img, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, time_weight)


def D_T_upwind(img, reg_z_over_reg=1.0, reg_time=0, mask_static=False, factor_reg_static=0, return_pytorch_tensor=False, time_weight=None):
    """Adjoint of D_upwind (tv_operators_GPU.py:828)."""
    return _adjoint("upwind", img, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, time_weight)


def D_T_central(img, reg_z_over_reg=1.0, reg_time=0, mask_static=False, factor_reg_static=0, return_pytorch_tensor=False, time_weight=None):
    """Adjoint of D_central (tv_operators_GPU.py:938); no Nz >= 5 / N >= 5 restriction."""
    return _adjoint("central", img, reg_z_over_reg, reg_time, mask_static, factor_reg_static, return_pytorch_tensor, time_weight)
