"""Seeded inputs shared by `make_golden.py` (reference side) and the tests (oracle / CUDA side)."""
import zlib

import numpy as np

SCHEMES = ("upwind", "downwind", "central", "hybrid")

# (Nz, M, N): 2-D, 3-D, 2-D+t, 4-D (even N, vector path), 4-D odd N (scalar path), M == 2 (centred time
# difference falls back to forward), Nz == 2, N == 8 with deeper z
SHAPES = [(1, 1, 5), (3, 1, 6), (1, 3, 5), (4, 3, 8), (3, 4, 7), (3, 2, 8), (2, 2, 5), (6, 1, 4)]

# (name, reg_z_over_reg, reg_time, use_mask_static, factor_reg_static)
WEIGHTS = [
    ("default", 1.0, 0.0, False, 0.0),
    ("noz", 0.0, 0.0, False, 0.0),
    ("zt", 0.5, 2.0 ** -5, False, 0.0),
    ("ztmask", 0.5, 2.0 ** -5, True, 4.0),
    ("unitmask", 1.0, 1.0, True, 0.25),
]


def _seed(*parts):
    return zlib.crc32("/".join(str(p) for p in parts).encode()) & 0x7FFFFFFF


def small_cases():
    out = []
    for (Nz, M, N) in SHAPES:
        for (wname, rz, rt, use_ms, fac) in WEIGHTS:
            for scheme in SCHEMES:
                if scheme == "central" and Nz == 2 and rz > 0:
                    continue  # the reference raises here: nothing to pin
                out.append(dict(key="%dx%dx%d/%s/%s" % (Nz, M, N, wname, scheme), shape=(Nz, M, N, N), scheme=scheme,
                                rz=rz, rt=rt, use_ms=use_ms, fac=fac, wname=wname))
    return out


def make_image(case, dtype=np.float64):
    rs = np.random.RandomState(_seed("x", case["shape"], case["wname"]))
    x = rs.rand(*case["shape"])
    # a flat patch so that zero gradient norms (0/0 := 0, inf in the returned norms) are exercised
    x[..., :2, :2] = 0.25
    return x.astype(dtype)


def make_field(case, shape, dtype=np.float64):
    rs = np.random.RandomState(_seed("p", case["shape"], case["wname"], case["scheme"]))
    return rs.randn(*shape).astype(dtype)


def make_mask_static(case):
    N = case["shape"][-1]
    rs = np.random.RandomState(_seed("ms", N))
    return rs.rand(1, 1, N, N) > 0.5


def weight_kwargs(case):
    kw = dict(reg_z_over_reg=case["rz"], reg_time=case["rt"])
    if case["use_ms"]:
        kw.update(mask_static=make_mask_static(case), factor_reg_static=case["fac"])
    return kw


# ---- known-answer inputs
def readme_volume():
    """README.md:78-80."""
    np.random.seed(0)
    return np.random.rand(20, 4, 100, 100)


def readme_mask_static():
    """SURVEY.md App. C."""
    return np.random.RandomState(1).rand(1, 1, 100, 100) > 0.5


def synthetic_image(N):
    """Piecewise-constant blocks plus a ramp, values in [0, 255], shape (1,1,N,N)."""
    ii, jj = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    img = 40.0 + 120.0 * ((ii // (N // 4) + jj // (N // 4)) % 2) + 60.0 * (jj / float(N))
    img[N // 3: N // 2, N // 3: N // 2] = 230.0
    return img.reshape(1, 1, N, N).astype(np.float64)


def cp_volume():
    rs = np.random.RandomState(7)
    base = np.zeros((4, 3, 8, 8))
    base[:, :, 2:6, 3:7] = 1.0
    base[2:, 1:, :, :4] += 0.5
    return base + 0.1 * rs.randn(*base.shape)


def cp_mask_static():
    return np.random.RandomState(8).rand(1, 1, 8, 8) > 0.5
