"""Marshalling between the reference's calling conventions (numpy arrays or torch tensors, any device) and
the device pointers the C ABI takes.  Mirrors tv_operators_GPU.py:92-131 (`type_like`) and :178-183."""
import ctypes

import numpy as np
import torch

from . import _lib

SCHEMES = ("upwind", "downwind", "central", "hybrid")


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("pytv_b200 needs a CUDA device: the library has no CPU path (use the reference's tv_CPU for that)")


def to_device(arr):
    """numpy / tensor -> contiguous CUDA tensor, float32 kept, everything else float64 (type_like).
    Returns (tensor, was_tensor)."""
    require_cuda()
    was_tensor = isinstance(arr, torch.Tensor)
    t = arr if was_tensor else torch.as_tensor(np.ascontiguousarray(arr))
    if t.dtype != torch.float32 and t.dtype != torch.float64:
        t = t.to(torch.float64)
    t = t.cuda()
    return t.contiguous(), was_tensor


def dtype_id(t):
    return _lib.F32 if t.dtype == torch.float32 else _lib.F64


def mask_static_to_device(mask_static, Ni, Nj):
    """bool `False` -> None; a (1,1,Ni,Nj) boolean array/tensor -> (Ni,Nj) uint8 CUDA tensor
    (tv_operators_GPU.py:236-240)."""
    if isinstance(mask_static, bool):
        return None
    m = mask_static if isinstance(mask_static, torch.Tensor) else torch.as_tensor(np.asarray(mask_static))
    m = m.reshape(-1, m.shape[-2], m.shape[-1])
    if m.shape[0] != 1 or m.shape[-2] != Ni or m.shape[-1] != Nj:
        raise RuntimeError("mask_static must have shape (1, 1, %d, %d), got %s" % (Ni, Nj, tuple(mask_static.shape)))
    return (m[0] != 0).to(torch.uint8).cuda().contiguous()


def image_shape(img):
    if img.ndim != 4:
        raise IndexError("pytv_b200 expects a 4-D image (Nz, M, N, N); got %d dimensions" % img.ndim)
    return tuple(int(s) for s in img.shape)


def time_scale_to_device(time_weight, shape, like):
    """EXTENSION (reference TODO, README.md:258): a (Nz, M, N, N) weight map of the time regularisation -> its square
    root as a contiguous device array of the compute dtype (what the C ABI takes as `time_scale`), or None."""
    if time_weight is None:
        return None
    w = time_weight if isinstance(time_weight, torch.Tensor) else torch.as_tensor(np.asarray(time_weight))
    w = torch.broadcast_to(w.to(like.device).to(torch.float64), tuple(shape))
    if bool((w < 0).any()):
        raise ValueError("time_weight must be >= 0")
    return torch.sqrt(w).to(like.dtype).contiguous()


def problem(scheme, t, shape, reg_z_over_reg, reg_time, factor_reg_static, ms, z_offset=0, Nz_global=None, ts=None):
    rz = float(reg_z_over_reg)
    return _lib.make_problem(scheme, dtype_id(t), shape, rz, float(reg_time), float(factor_reg_static),
                             ms.data_ptr() if ms is not None else None, z_offset, Nz_global, ts.data_ptr() if ts is not None else None)


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def reduce_workspace(pb, device):
    """Scratch of the calls that produce a scalar.  Zero-initialised: its first 64 bytes are the arrival counter of the reductions
    that finish inside the producing kernel (pytvb_tv, pytvb_gd_update), which the library leaves at zero after every call."""
    n = _lib.lib().pytvb_reduce_workspace_bytes(ctypes.byref(pb))
    return torch.zeros(n, dtype=torch.uint8, device=device)


def to_output(t, return_pytorch_tensor):
    return t if return_pytorch_tensor else t.detach().cpu().numpy()
