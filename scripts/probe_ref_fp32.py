#!/usr/bin/env python
"""float32 sub-gradient error of the REFERENCE's own torch GPU path (baseline/_ref/pytv/tv_GPU.py, unmodified) and of
this library, both against the float64 oracle on identical inputs (VERDICT r01 'pin the fp32 sub-gradient error').
Needs a CUDA device and baseline/_ref.  Writes gpurun_out/<tag>/ref_fp32.json.

    python scripts/probe_ref_fp32.py [--out gpurun_out/ref_fp32.json]
"""
import argparse
import importlib
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import cases  # noqa: E402
import pytv_b200 as pytv  # noqa: E402
from oracle import tv_oracle as orc  # noqa: E402

SCHEMES = ("upwind", "downwind", "central", "hybrid")


def load_ref():
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    warnings.filterwarnings("ignore")
    return importlib.import_module("pytv")


def stats(a, b):
    e = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))
    return {"max": float(e.max()), "p999": float(np.quantile(e, 0.999)), "mean": float(e.mean()), "frac_le_1e-5": float(np.mean(e <= 1e-5))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_fp32.json"))
    args = ap.parse_args()
    ref = load_ref()
    doc = {"gpu": torch.cuda.get_device_name(0), "cases": {}}
    vols = {"readme": cases.readme_volume().astype(np.float32)}
    rs = np.random.RandomState(5)
    vols["smooth64"] = (np.cumsum(rs.randn(6, 3, 64, 64), axis=-1) * 0.01).astype(np.float32)      # small gradients: ill-conditioned D/|D|
    vols["blocks"] = np.kron(rs.randint(0, 3, (4, 2, 8, 8)).astype(np.float32), np.ones((1, 1, 8, 8), np.float32)) + 0.001 * rs.rand(4, 2, 64, 64).astype(np.float32)
    for vname, x in vols.items():
        for wname, kw in (("default", {}), ("rt", dict(reg_time=2 ** -5)), ("rz_rt", dict(reg_z_over_reg=0.5, reg_time=1.0))):
            for scheme in SCHEMES:
                key = "%s/%s/%s" % (vname, wname, scheme)
                tv_o, G_o = orc.tv(x.astype(np.float64), scheme, **kw)
                r = {"tv_oracle": float(tv_o)}
                try:
                    tv_r, G_r = getattr(ref.tv_GPU, "tv_" + scheme)(x.copy(), **kw)
                    r["ref_gpu_vs_oracle"] = stats(G_r, G_o)
                    r["ref_gpu_tv_rel"] = abs(float(tv_r) - tv_o) / abs(tv_o)
                except Exception as e:
                    G_r = None
                    r["ref_gpu_error"] = repr(e)[:200]
                tv_m, G_m = getattr(pytv.tv_GPU, "tv_" + scheme)(x.copy(), **kw)
                r["ours_vs_oracle"] = stats(G_m, G_o)
                r["ours_tv_rel"] = abs(float(tv_m) - tv_o) / abs(tv_o)
                if G_r is not None:
                    r["ours_vs_ref_gpu"] = stats(G_m, G_r)
                # the oracle itself evaluated in float32 (numpy): the rounding floor of the formula
                _, G_o32 = orc.tv(x.copy(), scheme, **kw)
                r["oracle_f32_vs_oracle"] = stats(G_o32, G_o)
                doc["cases"][key] = r
                print(key, "ref %.2e  ours %.2e  oracle32 %.2e" % (r.get("ref_gpu_vs_oracle", {}).get("max", float("nan")), r["ours_vs_oracle"]["max"],
                                                                  r["oracle_f32_vs_oracle"]["max"]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(doc, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
