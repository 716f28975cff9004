#!/usr/bin/env python
"""Run every operator of the drop-in API a few times on one configuration (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytv_b200 as pytv  # noqa: E402

shape = tuple(int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (128, 4, 1024, 1024)
scheme = sys.argv[5] if len(sys.argv) > 5 else "hybrid"
kw = dict(reg_time=2 ** -5)
torch.manual_seed(0)
x = torch.rand(shape, device="cuda")
D = getattr(pytv.tv_operators_GPU, "D_" + scheme)
DT = getattr(pytv.tv_operators_GPU, "D_T_" + scheme)
tv = getattr(pytv.tv_GPU, "tv_" + scheme)
for _ in range(3):
    Dx = D(x, **kw)
    out = DT(Dx, **kw)
    l21 = pytv.tv_operators_GPU.compute_L21_norm(Dx)
    val, G = tv(x, return_pytorch_tensor=True, **kw)
    del Dx, out, G
torch.cuda.synchronize()
print("ok", float(val), float(l21))
