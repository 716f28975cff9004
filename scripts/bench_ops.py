#!/usr/bin/env python
"""Per-operator timings against the HBM roofline, and (when a copy of the reference travels in baseline/_ref)
the reference's own torch GPU path on the same box.  Writes one JSON document.

    python scripts/bench_ops.py [--out gpurun_out/ops.json] [--configs C2 C3 C4 C5] [--ref]
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
import pytv_b200 as pytv  # noqa: E402

CONFIGS = {
    "C2": ((20, 4, 100, 100), dict(reg_time=2 ** -5)),       # README volume
    "C2d": ((20, 4, 100, 100), dict()),                      # README volume, defaults (M is a batch axis)
    "C3": ((512, 1, 512, 512), dict()),
    "C4": ((128, 4, 1024, 1024), dict(reg_time=2 ** -5)),
    "C5": ((64, 8, 2048, 2048), dict(reg_time=2 ** -5)),
}
SCHEMES = ("upwind", "downwind", "central", "hybrid")


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


_flush = None


def timeit(fn, reps, flush):
    """median CUDA-event time in ms; L2 flushed between repetitions when the working set is small."""
    global _flush
    if flush and _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        if flush:
            _flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def ours(name, shape, kw, reps):
    V = int(np.prod(shape))
    flush = V * 4 * 10 < (1 << 30)
    torch.manual_seed(0)
    x = torch.rand(shape, device="cuda")
    out = {}
    for scheme in SCHEMES:
        D = getattr(pytv.tv_operators_GPU, "D_" + scheme)
        DT = getattr(pytv.tv_operators_GPU, "D_T_" + scheme)
        tv = getattr(pytv.tv_GPU, "tv_" + scheme)
        Dx = D(x, **kw)
        Nd = Dx.shape[1]
        r = {"Nd": Nd}
        for op, fn, bpv in (("D", lambda: D(x, **kw), 4 * (1 + Nd)), ("D_T", lambda: DT(Dx, **kw), 4 * (Nd + 1)),
                            ("L21", lambda: pytv.tv_operators_GPU.compute_L21_norm(Dx, return_pytorch_tensor=True), 4 * Nd),
                            ("tv", lambda: tv(x, return_pytorch_tensor=True, **kw), 8)):
            ms = timeit(fn, reps, flush)
            r[op] = {"ms": ms, "bytes_per_voxel": bpv, "GBps": bpv * V / ms / 1e6, "frac_of_measured_peak": bpv * V / ms / 1e6 / peak(),
                     "voxels_per_s": V / ms * 1e3}
        out[scheme] = r
        del Dx
        torch.cuda.empty_cache()
    return out


def reference(name, shape, kw, reps):
    """The unmodified reference torch path (pytv.tv_GPU / tv_operators_GPU) with CUDA tensors in."""
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import warnings
    warnings.filterwarnings("ignore")
    import importlib
    ref = importlib.import_module("pytv")
    V = int(np.prod(shape))
    torch.manual_seed(0)
    x = torch.rand(shape, device="cuda")
    out = {}
    for scheme in SCHEMES:
        D = getattr(ref.tv_operators_GPU, "D_" + scheme)
        DT = getattr(ref.tv_operators_GPU, "D_T_" + scheme)
        tv = getattr(ref.tv_GPU, "tv_" + scheme)
        r = {}
        try:
            Dx = D(x, **kw)
            for op, fn in (("D", lambda: D(x, **kw)), ("D_T", lambda: DT(Dx, **kw)),
                           ("L21", lambda: ref.tv_operators_GPU.compute_L21_norm(Dx)), ("tv", lambda: tv(x, return_pytorch_tensor=True, **kw))):
                try:
                    ms = timeit(fn, reps, False)
                    r[op] = {"ms": ms, "voxels_per_s": V / ms * 1e3}
                except Exception as e:   # e.g. D_T_central needs Nz >= 5
                    r[op] = {"error": repr(e)[:200]}
            del Dx
        except Exception as e:
            r["error"] = repr(e)[:200]
        out[scheme] = r
        torch.cuda.empty_cache()
    # README Chambolle-Pock iteration with device tensors (hybrid), the loop of README.md:145-157 with keepdims
    try:
        Dh, DTh = ref.tv_operators_GPU.D_hybrid, ref.tv_operators_GPU.D_T_hybrid
        x0 = x.clone()
        xc, yf = x.clone(), torch.zeros_like(x)
        ytv = torch.zeros_like(Dh(x, **kw))
        lam, sD, sA, tau = 0.1, 0.5, 1.0, 1.0 / 13.0

        def it():
            nonlocal xc, yf, ytv
            yf = (yf + sA * (xc - x0)) / (1.0 + sA)
            D_x = Dh(xc, **kw)
            pa = ytv + sD * D_x
            ytv = pa / torch.clamp(torch.sqrt(torch.sum(pa ** 2, dim=1, keepdim=True)) / lam, min=1.0)
            xc = xc - tau * yf - tau * DTh(ytv, **kw)
            return 0.5 * torch.sum((xc - x0) ** 2) + lam * ref.tv_operators_GPU.compute_L21_norm(D_x)
        ms = timeit(it, max(3, reps // 2), False)
        out["cp_readme_iteration_hybrid"] = {"ms": ms, "voxel_updates_per_s": V / ms * 1e3, "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    except Exception as e:
        out["cp_readme_iteration_hybrid"] = {"error": repr(e)[:300]}
    return out


def ours_cp(shape, kw, reps):
    torch.manual_seed(0)
    x0 = torch.rand(shape, device="cuda")
    V = int(np.prod(shape))
    res = {}
    for variant in ("readme", "rof"):
        s = pytv.CPSolver(x0, lam=0.1, scheme="hybrid", variant=variant, **kw)
        ms = timeit(lambda: s.step(1), reps, V * 4 * 12 < (1 << 30))
        Nd = s.Nd
        bpv = 4 * (3 * Nd + 5) if variant == "rof" else 4 * (3 * Nd + 6)
        res[variant] = {"ms": ms, "voxel_updates_per_s": V / ms * 1e3, "bytes_per_voxel": bpv, "GBps": bpv * V / ms / 1e6,
                        "frac_of_measured_peak": bpv * V / ms / 1e6 / peak()}
        del s
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ops.json"))
    ap.add_argument("--configs", nargs="+", default=["C2", "C2d", "C3", "C4"])
    ap.add_argument("--ref", action="store_true", help="also time the reference torch GPU path (needs baseline/_ref/pytv)")
    ap.add_argument("--ref-configs", nargs="+", default=["C2", "C2d", "C3"])
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    doc = {"gpu": torch.cuda.get_device_name(0), "hbm_peak_GBps": peak(), "ours": {}, "ours_cp_hybrid": {}, "reference_torch_gpu": {}}
    for c in args.configs:
        shape, kw = CONFIGS[c]
        doc["ours"][c] = {"shape": shape, "kw": {k: float(v) for k, v in kw.items()}, **ours(c, shape, kw, args.reps)}
        doc["ours_cp_hybrid"][c] = ours_cp(shape, kw, args.reps)
        torch.cuda.empty_cache()
    if args.ref:
        for c in args.ref_configs:
            shape, kw = CONFIGS[c]
            try:
                doc["reference_torch_gpu"][c] = {"shape": shape, **reference(c, shape, kw, max(3, args.reps // 2))}
            except Exception as e:
                doc["reference_torch_gpu"][c] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(doc, open(args.out, "w"), indent=1)
    for c, d in doc["ours"].items():
        for scheme in SCHEMES:
            r = d[scheme]
            print("%-4s %-9s Nd=%d  " % (c, scheme, r["Nd"]) + "  ".join("%s %.3f ms %.0f GB/s (%.2f)" % (op, r[op]["ms"], r[op]["GBps"], r[op]["frac_of_measured_peak"])
                                                                      for op in ("D", "D_T", "L21", "tv")))
        print(c, "CP hybrid:", {k: "%.3f ms %.2f" % (v["ms"], v["frac_of_measured_peak"]) for k, v in doc["ours_cp_hybrid"][c].items()})
    for c, d in doc["reference_torch_gpu"].items():
        print("REF", c, json.dumps({k: ({o: (round(v[o]["ms"], 3) if "ms" in v[o] else "err") for o in v} if k in SCHEMES else v) for k, v in d.items() if k != "shape"})[:900])


if __name__ == "__main__":
    main()
