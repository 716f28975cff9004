#!/usr/bin/env python
"""Latency-bound configurations (BASELINE configs 1 and 2): device-resident times of a tv_<scheme> call (TVPlan graph replay), of
the README's sub-gradient descent and Chambolle-Pock iterations on a cameraman-sized image, and the literal README loops
through the numpy drop-in API.  PYTVB_TV_PATH=sweeps|tile selects the tv implementation (read once per process).

    python scripts/bench_small.py [--out gpurun_out/small.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import pytv_b200 as pytv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "small.json"))
args = ap.parse_args()


def ev_time(fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {"tv_path": os.environ.get("PYTVB_TV_PATH", "auto")}
# ---- (1) one tv_<scheme> call, device-resident: graph replay of the plan, per scheme
for name, shape, kw in (("C2_readme_volume_rt", (20, 4, 100, 100), dict(reg_time=2 ** -5)), ("C2_readme_volume_default", (20, 4, 100, 100), dict()),
                        ("C1_256", (1, 1, 256, 256), dict()), ("C1_512", (1, 1, 512, 512), dict())):
    r = {}
    for scheme in ("hybrid", "upwind", "central"):
        for dt in (torch.float32, torch.float64):
            plan = pytv.TVPlan(scheme, shape, dt, **kw)
            plan.x.copy_(torch.rand(shape, dtype=dt, device="cuda"))
            r["%s_%s_us" % (scheme, "f32" if dt == torch.float32 else "f64")] = 1e3 * ev_time(plan.run, 200)
    out["tv_plan_replay " + name] = r
# ---- (2) the README loops on a cameraman-sized image
for N in (256, 512):
    ii, jj = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    img = (40.0 + 120.0 * ((ii // (N // 4) + jj // (N // 4)) % 2) + 60.0 * (jj / float(N))).reshape(1, 1, N, N)
    noisy = img + 100 * np.random.RandomState(0).rand(*img.shape)
    r = {}
    # (a) README sub-gradient descent loop, numpy in / numpy out through tv_GPU.tv_hybrid (README.md:107-124)
    x = noisy.copy()
    pytv.tv_GPU.tv_hybrid(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(300):
        tv, G = pytv.tv_GPU.tv_hybrid(x)
        x += -5e-3 * ((x - noisy) + 25.0 * G)
    r["gd_readme_loop_numpy_api_ms_per_it"] = (time.perf_counter() - t0) / 300 * 1e3
    # (a') the same loop resident on the device: gd_denoise (tv kernel + fused update, the 300 iterations in one CUDA graph)
    for dtype in (np.float64, np.float32):
        nd = torch.as_tensor(noisy.astype(dtype)).cuda()
        pytv.gd_denoise(nd, 25.0, 300, 5e-3)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pytv.gd_denoise(nd, 25.0, 300, 5e-3, return_losses=True)
        torch.cuda.synchronize()
        r["gd_denoise_call_%s_us_per_it_incl_capture" % np.dtype(dtype).name] = (time.perf_counter() - t0) / 300 * 1e6
        # the replay alone: a plan + the fused update captured once, replayed
        plan = pytv.TVPlan("hybrid", (1, 1, N, N), torch.float32 if dtype == np.float32 else torch.float64, graph=False)
        plan.x.copy_(nd)
        import ctypes
        from pytv_b200 import _dev, _lib
        lib = _lib.lib()
        fid = torch.zeros(1, dtype=torch.float64, device="cuda")

        def it():
            plan.run()
            _lib.check(lib.pytvb_gd_update(ctypes.byref(plan.pb), _dev.ptr(plan.x), _dev.ptr(nd), _dev.ptr(plan.G), 5e-3, 25.0, _dev.ptr(fid),
                                           _dev.ptr(plan._ws_r), _dev.stream_ptr()))
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            it()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(50):
                it()
        r["gd_iteration_device_%s_us" % np.dtype(dtype).name] = 1e3 * ev_time(g.replay, 6) / 50
    # (b) fused CP solver, eager and graph
    for dtype in (np.float64, np.float32):
        s = pytv.CPSolver(noisy.astype(dtype), lam=25.0, scheme="hybrid", variant="readme", tau=1 / 9.0)
        s.step(10); torch.cuda.synchronize(); t0 = time.perf_counter()
        s.step(300); torch.cuda.synchronize()
        r["cp_eager_%s_us_per_it" % np.dtype(dtype).name] = (time.perf_counter() - t0) / 300 * 1e6
        s.capture_graph(iterations=10)
        s.step(10)
        r["cp_graph_%s_us_per_it" % np.dtype(dtype).name] = 1e3 * ev_time(lambda: s.step(300), 1) / 300
    out["%dx%d" % (N, N)] = r
print(json.dumps(out, indent=1))
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(out, open(args.out, "w"), indent=1)
