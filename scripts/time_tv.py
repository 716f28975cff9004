import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytv_b200 as pytv
shape = (128, 4, 1024, 1024)
torch.manual_seed(0)
x = torch.rand(shape, device="cuda")
for scheme in (sys.argv[1:] or ("hybrid", "upwind", "central")):
    f = getattr(pytv.tv_GPU, "tv_" + scheme)
    for _ in range(3): f(x, return_pytorch_tensor=True, reg_time=2**-5)
    ts = []
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(x, return_pytorch_tensor=True, reg_time=2**-5); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(os.environ.get("PYTVB_LIB_PATH", "default").split("_")[-1], scheme, "tv %.3f ms" % np.median(ts))
