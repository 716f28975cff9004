import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytv_b200 as pytv
shape = (128, 4, 1024, 1024)
torch.manual_seed(0)
x = torch.rand(shape, device="cuda")
for scheme in ("hybrid", "upwind"):
    f = getattr(pytv.tv_operators_GPU, "D_" + scheme)
    for _ in range(3): out = f(x, reg_time=2**-5)
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = f(x, reg_time=2**-5); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    V = x.numel(); Nd = out.shape[1]
    print(os.environ.get("PYTVB_LIB_PATH", "default").split("_")[-1], scheme, "D %.3f ms  %.0f GB/s" % (np.median(ts), 4 * (1 + Nd) * V / np.median(ts) / 1e6))
