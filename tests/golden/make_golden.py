"""Generate the golden fixtures in this directory from the UNMODIFIED reference (PyTV-4D v1.1.2).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports `pytv.tv_CPU` / `pytv.tv_operators_CPU` from /root/reference, runs them on small seeded
inputs and stores the outputs.  Inputs are *not* stored: `cases.py` regenerates them from the same
`np.random.RandomState` seeds (the legacy generator is stable across numpy versions).  The tests never
import the reference; they read these files.

Files written:
  golden_small.npz   reference D, D_T, TV, G, grad-norms for every (shape, weights, scheme) in cases.py
  golden_kat.json    scalars: README volume known answers, 5x5 delta images, CP/GD loss sequences
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")
sys.path.insert(0, os.environ.get("PYTV_REFERENCE", "/root/reference"))

import cases  # noqa: E402
import pytv   # noqa: E402  (the reference)

opC = pytv.tv_operators_CPU
tvC = pytv.tv_CPU


def ref_D(scheme):
    return getattr(opC, "D_" + scheme)


def ref_DT(scheme):
    return getattr(opC, "D_T_" + scheme)


def ref_tv(scheme):
    return getattr(tvC, "tv_" + scheme)


def small_cases():
    out = {}
    for case in cases.small_cases():
        key = case["key"]
        x = cases.make_image(case)
        kw = cases.weight_kwargs(case)
        scheme = case["scheme"]
        Dx = ref_D(scheme)(x.copy(), **kw)
        p = cases.make_field(case, Dx.shape)
        DTp = ref_DT(scheme)(p.copy(), **kw)
        tv, G, norms = ref_tv(scheme)(x.copy(), return_grad_norms=True, **kw)
        out[key + "/D"] = Dx
        out[key + "/DT"] = DTp
        out[key + "/tv"] = np.float64(tv)
        out[key + "/G"] = G
        out[key + "/norms"] = norms
    return out


def readme_cp_loop(x0, nb_it, lam, scheme, sigma_D=0.5, sigma_A=1.0, tau=1.0 / 9.0, **kw):
    """README.md:139-158 with the projection's norm kept broadcastable (keepdims) so that Nz > 1 works."""
    D, DT = ref_D(scheme), ref_DT(scheme)
    x = np.copy(x0)
    y_f = np.zeros_like(x0)
    y_tv = np.zeros_like(D(x0, **kw))
    losses = []
    for _ in range(nb_it):
        y_f = (y_f + sigma_A * (x - x0)) / (1.0 + sigma_A)
        D_x = D(x, **kw)
        prox_argument = y_tv + sigma_D * D_x
        y_tv = prox_argument / np.maximum(1.0, np.sqrt(np.sum(prox_argument ** 2, axis=1, keepdims=True)) / lam)
        x = x - tau * y_f - tau * DT(y_tv, **kw)
        losses.append(float(0.5 * np.sum(np.square(x - x0)) + lam * opC.compute_L21_norm(D_x)))
    return x, y_f, y_tv, losses


def rof_cp_loop(x0, nb_it, lam, scheme, sigma, tau, theta=1.0, **kw):
    """Chambolle-Pock Alg. 1 for 0.5||x-x0||^2 + lam TV(x), built from the reference operators."""
    D, DT = ref_D(scheme), ref_DT(scheme)
    x = np.copy(x0)
    xbar = np.copy(x0)
    y = np.zeros_like(D(x0, **kw))
    energies = []
    for _ in range(nb_it):
        Dxb = D(xbar, **kw)
        pa = y + sigma * Dxb
        y = pa / np.maximum(1.0, np.sqrt(np.sum(pa ** 2, axis=1, keepdims=True)) / lam)
        x_new = (x - tau * DT(y, **kw) + tau * x0) / (1.0 + tau)
        xbar = x_new + theta * (x_new - x)
        x = x_new
        energies.append(float(0.5 * np.sum(np.square(x - x0)) + lam * opC.compute_L21_norm(Dxb)))
    return x, xbar, y, energies


def gd_loop(x0, nb_it, lam, step, scheme):
    """README.md:107-124 with tv_CPU."""
    x = np.copy(x0)
    losses, tv = [], 0.0
    for _ in range(nb_it):
        tv, G = ref_tv(scheme)(x)
        x += -step * ((x - x0) + lam * G)
        losses.append(float(0.5 * np.sum(np.square(x - x0)) + lam * tv))
    return x, losses, float(tv)


def kats():
    kat = {}
    # --- README volume (README.md:76-93), SURVEY App. C
    img = cases.readme_volume()
    ms = cases.readme_mask_static()
    vol = {}
    for scheme in cases.SCHEMES:
        e = {}
        tv, G = ref_tv(scheme)(img.copy())
        e["default"] = dict(tv=float(tv), sum_abs_G=float(np.abs(G).sum()), G_3_1_5_7=float(G[3, 1, 5, 7]))
        tv, G = ref_tv(scheme)(img.copy(), reg_time=2 ** -5)
        e["rt"] = dict(tv=float(tv), sum_abs_G=float(np.abs(G).sum()), G_3_1_5_7=float(G[3, 1, 5, 7]))
        kw = dict(reg_z_over_reg=0.5, reg_time=2 ** -5, mask_static=ms, factor_reg_static=4.0)
        Dx = ref_D(scheme)(img.copy(), **kw)
        DTD = ref_DT(scheme)(Dx, **kw)
        tv, G = ref_tv(scheme)(img.copy(), **kw)
        e["weighted"] = dict(Nd=int(Dx.shape[1]), tv=float(tv), l21=float(opC.compute_L21_norm(Dx)),
                             sum_abs_D=float(np.abs(Dx).sum()), sum_abs_DTD=float(np.abs(DTD).sum()),
                             sum_abs_G=float(np.abs(G).sum()))
        tv32, G32 = ref_tv(scheme)(img.astype(np.float32))
        e["float32_input"] = dict(tv=float(tv32))
        vol[scheme] = e
    kat["readme_volume"] = vol

    # --- 5x5 delta image (examples/b_TV_discretizations_math.ipynb)
    delta = {}
    for scheme in cases.SCHEMES:
        A = np.zeros((1, 1, 5, 5))
        A[0, 0, 2, 2] = 1.0
        tv, G = ref_tv(scheme)(A)
        delta[scheme] = dict(tv=float(tv), G=G[0, 0].tolist())
    kat["delta5"] = delta

    # --- denoising loops on a synthetic 64x64 image (the cameraman asset is not redistributed here)
    x_true = cases.synthetic_image(64)
    rs = np.random.RandomState(0)
    noisy = x_true + 100 * rs.rand(*x_true.shape)
    x, losses, tv = gd_loop(noisy, 50, 25.0, 5e-3, "hybrid")
    kat["gd_synthetic64"] = dict(loss_first=losses[0], loss_last=losses[-1], tv_last=tv, sum_x=float(x.sum()),
                                 losses=losses)
    x, y_f, y_tv, losses = readme_cp_loop(noisy, 50, 25.0, "hybrid")
    kat["cp_readme_synthetic64"] = dict(loss_first=losses[0], loss_last=losses[-1], sum_x=float(x.sum()),
                                        sum_abs_ytv=float(np.abs(y_tv).sum()), losses=losses)

    # --- CP loops on a small 4-D volume, all schemes, weights on
    cp4 = {}
    x0 = cases.cp_volume()
    ms4 = cases.cp_mask_static()
    kw = dict(reg_z_over_reg=0.5, reg_time=2 ** -5, mask_static=ms4, factor_reg_static=4.0)
    for scheme in cases.SCHEMES:
        x, y_f, y_tv, losses = readme_cp_loop(x0, 10, 0.2, scheme, tau=1.0 / 17.0, **kw)
        xr, xbar, y, energies = rof_cp_loop(x0, 10, 0.2, scheme, sigma=0.5, tau=1.0 / 17.0, **kw)
        cp4[scheme] = dict(readme_losses=losses, readme_sum_x=float(x.sum()), readme_sum_abs_y=float(np.abs(y_tv).sum()),
                           readme_x_probe=float(x[1, 1, 3, 4]),
                           rof_energies=energies, rof_sum_x=float(xr.sum()), rof_sum_abs_y=float(np.abs(y).sum()),
                           rof_sum_xbar=float(xbar.sum()), rof_x_probe=float(xr[1, 1, 3, 4]))
    kat["cp_small4d"] = cp4

    # --- cameraman (README.md:107-158), only scalars; needs the reference asset
    cam = pytv.utils.cameraman().astype(np.float64)
    cam = cam.reshape((1, 1) + cam.shape)
    np.random.seed(0)
    noisy = cam + 100 * np.random.rand(*cam.shape)
    x, losses, tv = gd_loop(noisy, 300, 25.0, 5e-3, "hybrid")
    kat["gd_cameraman"] = dict(loss_first=losses[0], loss_last=losses[-1], tv_last=tv, sum_x=float(x.sum()))
    x, y_f, y_tv, losses = readme_cp_loop(noisy, 300, 25.0, "hybrid")
    kat["cp_readme_cameraman"] = dict(loss_first=losses[0], loss_last=losses[-1], sum_x=float(x.sum()))
    kat["cameraman_checksum"] = dict(shape=list(cam.shape), sum=float(cam.sum()), sum_sq=float((cam * cam).sum()))
    return kat


if __name__ == "__main__":
    small = small_cases()
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **small)
    with open(os.path.join(HERE, "golden_kat.json"), "w") as f:
        json.dump(kats(), f, indent=1)
    print("wrote %d arrays" % len(small))
