#!/usr/bin/env python
"""Latency-bound configurations: BASELINE config 1 (cameraman-sized 256x256 / 512x512 image, 300 iterations of the
README loops) - eager launches vs CUDA-graph replay, and the literal README loops through the drop-in API."""
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

out = {}
for N in (256, 512):
    ii, jj = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    img = (40.0 + 120.0 * ((ii // (N // 4) + jj // (N // 4)) % 2) + 60.0 * (jj / float(N))).reshape(1, 1, N, N)
    noisy = img + 100 * np.random.RandomState(0).rand(*img.shape)
    r = {}
    # (a) README sub-gradient descent loop, numpy in / numpy out through tv_GPU.tv_hybrid (README.md:107-124)
    x = noisy.copy()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(300):
        tv, G = pytv.tv_GPU.tv_hybrid(x)
        x += -5e-3 * ((x - noisy) + 25.0 * G)
    r["gd_readme_loop_numpy_api_ms_per_it"] = (time.perf_counter() - t0) / 300 * 1e3
    # (b) fused CP solver, eager and graph
    for dtype in (np.float64, np.float32):
        s = pytv.CPSolver(noisy.astype(dtype), lam=25.0, scheme="hybrid", variant="readme", tau=1 / 9.0)
        s.step(10); torch.cuda.synchronize(); t0 = time.perf_counter()
        s.step(300); torch.cuda.synchronize()
        r["cp_eager_%s_ms_per_it" % np.dtype(dtype).name] = (time.perf_counter() - t0) / 300 * 1e3
        s.capture_graph(iterations=10)
        s.step(10); torch.cuda.synchronize(); t0 = time.perf_counter()
        s.step(300); torch.cuda.synchronize()
        r["cp_graph_%s_ms_per_it" % np.dtype(dtype).name] = (time.perf_counter() - t0) / 300 * 1e3
    out["%dx%d" % (N, N)] = r
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "small.json"), "w"), indent=1)
