"""CUDA-event time of tv_<scheme> on the C4 slab (or --shape Nz M Ni Nj), through the drop-in call with a CUDA tensor in.
    python scripts/time_tv.py [schemes ...] [--shape 128 4 1024 1024] [--rt 0.03125]
PYTVB_TV_PATH=sweeps times the two-sweep fallback instead of the single-sweep tile kernel."""
import argparse, os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytv_b200 as pytv
ap = argparse.ArgumentParser()
ap.add_argument("schemes", nargs="*", default=["hybrid", "upwind", "downwind", "central"])
ap.add_argument("--shape", type=int, nargs=4, default=[128, 4, 1024, 1024])
ap.add_argument("--rt", type=float, default=2 ** -5)
ap.add_argument("--reps", type=int, default=8)
args = ap.parse_args()
shape = tuple(args.shape)
torch.manual_seed(0)
x = torch.rand(shape, device="cuda")
V = x.numel()
for scheme in args.schemes:
    f = getattr(pytv.tv_GPU, "tv_" + scheme)
    for _ in range(3): f(x, return_pytorch_tensor=True, reg_time=args.rt)
    ts = []
    for _ in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(x, return_pytorch_tensor=True, reg_time=args.rt); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    print(os.environ.get("PYTVB_TV_PATH", "tile"), "x".join(map(str, shape)), scheme, "tv %.3f ms  %.0f GB/s of the 8 B/voxel" % (ms, 8 * V / ms / 1e6))
