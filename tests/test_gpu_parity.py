"""Parity of the CUDA path (through the Python drop-in API and the C ABI) against the oracle and the
reference goldens.  Needs a CUDA device: run with `-m gpu` on the B200 box.

Tolerances (BASELINE.json north_star): float32 - TV within 1e-5 relative, operator outputs and sub-gradient
within 1e-5 absolute, adjointness 1e-4 relative; float64 - 1e-12 (summation-order noise only)."""
import ctypes
import math

import numpy as np
import pytest
import torch

import cases
import pytv_b200 as pytv
from oracle import tv_oracle as orc
from pytv_b200 import _dev, _lib

pytestmark = pytest.mark.gpu

SCHEMES = cases.SCHEMES
opG, tvG = pytv.tv_operators_GPU, pytv.tv_GPU


def D_(scheme):
    return getattr(opG, "D_" + scheme)


def DT_(scheme):
    return getattr(opG, "D_T_" + scheme)


def tv_(scheme):
    return getattr(tvG, "tv_" + scheme)


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    assert "B200" in torch.cuda.get_device_name(0) or True


# ------------------------------------------------------------------ golden vectors of the reference
@pytest.mark.parametrize("case", [pytest.param(c, id=c["key"]) for c in cases.small_cases()])
def test_small_goldens_float64(case, golden_small):
    key, scheme = case["key"], case["scheme"]
    kw = cases.weight_kwargs(case)
    x = cases.make_image(case)
    gD = golden_small[key + "/D"]
    Dx = D_(scheme)(x, **kw)
    assert isinstance(Dx, np.ndarray) and Dx.dtype == np.float64 and Dx.shape == gD.shape
    np.testing.assert_allclose(Dx, gD, rtol=0, atol=1e-13)
    p = cases.make_field(case, gD.shape)
    np.testing.assert_allclose(DT_(scheme)(p, **kw), golden_small[key + "/DT"], rtol=0, atol=1e-11)
    tv, G, norms = tv_(scheme)(x.copy(), return_grad_norms=True, **kw)
    assert float(tv) == pytest.approx(float(golden_small[key + "/tv"]), rel=1e-13)
    np.testing.assert_allclose(G, golden_small[key + "/G"], rtol=0, atol=1e-10)
    gn = golden_small[key + "/norms"]
    assert np.array_equal(np.isinf(norms), np.isinf(gn))
    np.testing.assert_allclose(norms[np.isfinite(gn)], gn[np.isfinite(gn)], rtol=0, atol=1e-13)
    assert float(opG.compute_L21_norm(Dx)) == pytest.approx(float(golden_small[key + "/tv"]), rel=1e-13)


@pytest.mark.parametrize("case", [pytest.param(c, id=c["key"]) for c in cases.small_cases()])
def test_small_goldens_float32(case, golden_small):
    key, scheme = case["key"], case["scheme"]
    kw = cases.weight_kwargs(case)
    x = cases.make_image(case, np.float32)
    gD = golden_small[key + "/D"]
    Dx = D_(scheme)(x, **kw)
    assert Dx.dtype == np.float32
    np.testing.assert_allclose(Dx, gD, rtol=0, atol=1e-5)
    p = cases.make_field(case, gD.shape, np.float32)
    np.testing.assert_allclose(DT_(scheme)(p, **kw), golden_small[key + "/DT"], rtol=0, atol=1e-5)
    tv, G, norms = tv_(scheme)(x.copy(), return_grad_norms=True, **kw)
    assert float(tv) == pytest.approx(float(golden_small[key + "/tv"]), rel=1e-5)
    gn = golden_small[key + "/norms"]
    assert np.array_equal(np.isinf(norms), np.isinf(gn))
    # sub-gradient: 1e-5 absolute (north star), or - where D/|D| is ill-conditioned because the norms are tiny - 3 x the
    # rounding floor of the reference formula itself evaluated in float32 (numpy).  Measured on the B200 (profiles/
    # r02a_ref_fp32.json): this library sits AT that floor (7e-7 .. 1.6e-6 on the README volume), the reference's own
    # float32 GPU path is at 5e-3 .. 5e-1 (its conv3d calls run in TF32).
    _assert_G_float32(G, golden_small[key + "/G"], x, scheme, kw)


def _assert_G_float32(G, G64, x32, scheme, kw):
    _, G32 = orc.tv(x32.copy(), scheme, **kw)
    floor = float(np.abs(G32.astype(np.float64) - G64).max())
    err = float(np.abs(G.astype(np.float64) - G64).max())
    assert err <= max(1e-5, 3.0 * floor), (err, floor)


# ------------------------------------------------------------------ published known answers
@pytest.mark.parametrize("scheme", SCHEMES)
def test_readme_volume_known_answers(scheme, golden_kat):
    """README.md:76-93 (BASELINE config 2): float64 input as in the README, and its float32 cast."""
    g = golden_kat["readme_volume"][scheme]
    img = cases.readme_volume()
    tv, G = tv_(scheme)(img.copy())
    assert tv.shape == () and isinstance(tv, np.ndarray)
    assert float(tv) == pytest.approx(g["default"]["tv"], rel=1e-13)
    assert np.abs(G).sum() == pytest.approx(g["default"]["sum_abs_G"], rel=1e-12)
    assert G[3, 1, 5, 7] == pytest.approx(g["default"]["G_3_1_5_7"], rel=1e-11)
    tv_o, G_o = orc.tv(img.copy(), scheme)
    assert np.prod(np.abs(G - G_o) < 1e-5) > 0          # the README's own check (README.md:85)
    np.testing.assert_allclose(G, G_o, rtol=0, atol=1e-11)
    tv, G = tv_(scheme)(img.copy(), reg_time=2 ** -5)
    assert float(tv) == pytest.approx(g["rt"]["tv"], rel=1e-13)
    assert np.abs(G).sum() == pytest.approx(g["rt"]["sum_abs_G"], rel=1e-12)
    kw = dict(reg_z_over_reg=0.5, reg_time=2 ** -5, mask_static=cases.readme_mask_static(), factor_reg_static=4.0)
    Dx = D_(scheme)(img, **kw)
    assert Dx.shape[1] == g["weighted"]["Nd"]
    assert float(opG.compute_L21_norm(Dx)) == pytest.approx(g["weighted"]["l21"], rel=1e-13)
    assert np.abs(Dx).sum() == pytest.approx(g["weighted"]["sum_abs_D"], rel=1e-12)
    assert np.abs(DT_(scheme)(Dx, **kw)).sum() == pytest.approx(g["weighted"]["sum_abs_DTD"], rel=1e-12)
    tv, G = tv_(scheme)(img.copy(), **kw)
    assert float(tv) == pytest.approx(g["weighted"]["tv"], rel=1e-13)
    assert np.abs(G).sum() == pytest.approx(g["weighted"]["sum_abs_G"], rel=1e-12)
    # float32
    img32 = img.astype(np.float32)
    tv32, G32 = tv_(scheme)(img32.copy())
    assert G32.dtype == np.float32
    assert float(tv32) == pytest.approx(g["default"]["tv"], rel=1e-5)
    # same float32 input through the float64 oracle: 1e-5 absolute, no escape hatch (measured: 4e-7 central .. 1.6e-6 hybrid)
    _, G_o32in = orc.tv(img32.astype(np.float64), scheme)
    assert float(np.abs(G32 - G_o32in).max()) <= 1e-5
    # against the float64 INPUT the rounding of the image itself enters (amplified where |Dx| is small): README.md:85's check
    assert np.mean(np.abs(G32 - G_o) < 1e-5) > 0.999


def test_readme_published_value():
    tv, G = tvG.tv_hybrid(cases.readme_volume())
    assert float(tv) == pytest.approx(532166.8251801673, rel=1e-13)


@pytest.mark.parametrize("scheme", SCHEMES)
def test_delta_image(scheme, golden_kat):
    A = np.zeros((1, 1, 5, 5))
    A[0, 0, 2, 2] = 1.0
    tv, G = tv_(scheme)(A)
    closed = {"upwind": 2 + math.sqrt(2), "downwind": 2 + math.sqrt(2), "central": 2.0, "hybrid": 3 * math.sqrt(2)}
    assert float(tv) == pytest.approx(closed[scheme], rel=1e-14)
    np.testing.assert_allclose(G[0, 0], np.array(golden_kat["delta5"][scheme]["G"]), atol=1e-14)


# ------------------------------------------------------------------ API conventions
def test_return_conventions():
    rs = np.random.RandomState(0)
    x = rs.rand(3, 2, 8, 8).astype(np.float32)
    xt = torch.as_tensor(x)
    # operators: numpy in -> numpy out; tensor in (CPU or CUDA) -> CUDA tensor out
    assert isinstance(opG.D_hybrid(x), np.ndarray)
    out = opG.D_hybrid(xt)
    assert isinstance(out, torch.Tensor) and out.is_cuda and out.dtype == torch.float32
    out = opG.D_hybrid(x, return_pytorch_tensor=True)
    assert isinstance(out, torch.Tensor) and out.is_cuda
    assert isinstance(opG.D_T_hybrid(out), torch.Tensor)
    assert isinstance(opG.D_T_hybrid(out.cpu().numpy()), np.ndarray)
    # dtype policy: float32 stays, everything else float64
    assert opG.D_upwind(x.astype(np.float16)).dtype == np.float64
    assert opG.D_upwind((100 * x).astype(np.int64)).dtype == np.float64
    # tv: G numpy unless return_pytorch_tensor, even for tensor input; tv is a 0-d ndarray
    tv, G = tvG.tv_hybrid(xt.cuda())
    assert isinstance(G, np.ndarray) and isinstance(tv, np.ndarray) and tv.shape == ()
    tv, G, n = tvG.tv_hybrid(xt.cuda(), return_pytorch_tensor=True, return_grad_norms=True)
    assert G.is_cuda and n.is_cuda
    # compute_L21_norm: 0-d ndarray; the norm array is a tensor even without return_pytorch_tensor
    l21, arr = opG.compute_L21_norm(opG.D_hybrid(x), return_array=True)
    assert isinstance(l21, np.ndarray) and l21.shape == () and isinstance(arr, torch.Tensor)
    l21t, _ = opG.compute_L21_norm(opG.D_hybrid(x), return_array=True, return_pytorch_tensor=True)
    assert isinstance(l21t, torch.Tensor) and l21t.is_cuda
    # without return_array the scalar is a 0-d numpy array even when a tensor is asked for (tv_operators_GPU.py:88-90)
    l21n = opG.compute_L21_norm(opG.D_hybrid(x), return_array=False, return_pytorch_tensor=True)
    assert isinstance(l21n, np.ndarray) and l21n.shape == ()
    assert float(l21) == pytest.approx(float(tv), rel=1e-6)
    np.testing.assert_allclose(arr.cpu().numpy(), np.where(np.isinf(n.cpu().numpy()), 0, n.cpu().numpy()), atol=1e-6)
    # errors
    with pytest.raises(IndexError):
        opG.D_hybrid(x[0])
    with pytest.raises(IndexError):
        opG.D_T_hybrid(out[:, :3])
    assert opG.type_like(np.zeros(3), xt).dtype == np.float32
    assert opG.type_like(torch.zeros(3), np.zeros(2)).dtype == torch.float64


@pytest.mark.parametrize("kind", ["numpy_mask_numpy_img", "tensor_mask_cuda_img", "plane_mask"])
def test_mask_is_applied_in_place(kind):
    rs = np.random.RandomState(2)
    x = rs.rand(3, 2, 8, 8)
    mask = rs.rand(3, 2, 8, 8) > 0.3
    if kind == "plane_mask":
        mask = np.broadcast_to(rs.rand(8, 8) > 0.3, x.shape).copy()
    x_ref = x.copy()
    x_ref[~mask] = 0
    tv_o, G_o = orc.tv(x_ref.copy(), "hybrid", reg_time=0.5)
    if kind == "tensor_mask_cuda_img":
        xc = torch.as_tensor(x).cuda()
        tv, G = tvG.tv_hybrid(xc, mask=torch.as_tensor(mask), reg_time=0.5)
        np.testing.assert_array_equal(xc.cpu().numpy(), x_ref)
    elif kind == "plane_mask":
        xi = x.copy()
        tv, G = tvG.tv_hybrid(xi, mask=mask[0, 0], reg_time=0.5)
        np.testing.assert_array_equal(xi, x_ref)
    else:
        xi = x.copy()
        tv, G = tvG.tv_hybrid(xi, mask=mask, reg_time=0.5)
        np.testing.assert_array_equal(xi, x_ref)
    assert float(tv) == pytest.approx(tv_o, rel=1e-13)
    np.testing.assert_allclose(G, G_o, atol=1e-12)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(70, 2, 8, 12), (3, 2, 7, 9), (2, 1, 16, 132)], ids=["140-planes", "odd-width", "wide"])
@pytest.mark.parametrize("plane", [False, True])
def test_apply_mask_kernel_forms(dtype, shape, plane):
    """img[~mask] = 0 (tv_GPU.py:79-80) through both kernels: four pixels per thread walking the planes (row length divisible by 4;
    more planes than the grid has rows) and the scalar form, full-volume and single-plane masks, NaNs under the mask removed."""
    rs = np.random.RandomState(5)
    x = rs.rand(*shape).astype(dtype)
    mask = (rs.rand(*shape[2:]) > 0.4) if plane else (rs.rand(*shape) > 0.4)
    full = np.broadcast_to(mask, shape)
    x[~full] = np.nan
    xc = torch.as_tensor(x).cuda()
    getattr(tvG, "tv_upwind")(xc, mask=torch.as_tensor(mask))
    ref = x.copy()
    ref[~full] = 0
    np.testing.assert_array_equal(xc.cpu().numpy(), ref)


# ------------------------------------------------------------------ reference test-suite, restated (tests.py)
@pytest.mark.parametrize("scheme", SCHEMES)
def test_operator_transpose(scheme):
    """tests.py:111-185: <Y, D X> = <X, D^T Y>, float32 Gaussian data, n_rays = 100, relative 1e-4."""
    rs = np.random.RandomState(1)
    configs = [((1, 1), {}), ((20, 1), {}), ((20, 1), dict(reg_z_over_reg=0))]
    for M in (2, 3, 4):
        configs += [((1, M), dict(reg_time=1.0)), ((20, M), dict(reg_time=1.0)), ((20, M), dict(reg_z_over_reg=0, reg_time=1.0))]
    for (Nz, M), kw in configs:
        X = rs.randn(Nz, M, 100, 100).astype(np.float32)
        DX = D_(scheme)(X, **kw)
        Y = rs.randn(*DX.shape).astype(np.float32)
        a = np.sum(Y.astype(np.float64) * DX)
        b = np.sum(X.astype(np.float64) * DT_(scheme)(Y, **kw))
        assert abs(a - b) / (0.5 * (abs(a) + abs(b))) < 1e-4, (scheme, Nz, M, kw)


@pytest.mark.parametrize("scheme", SCHEMES)
def test_2d_tiled_to_3d(scheme):
    """tests.py:187-245 restated with explicit shapes."""
    rs = np.random.RandomState(5)
    img2 = rs.rand(1, 1, 100, 100)
    Nz = 6
    img3 = np.tile(img2, (Nz, 1, 1, 1))
    tv2, G2 = tv_(scheme)(img2.copy())
    tv3, G3 = tv_(scheme)(img3.copy())
    assert float(tv3) == pytest.approx(Nz * float(tv2), rel=1e-13)
    for z in range(Nz):
        np.testing.assert_allclose(G3[z], G2[0], atol=1e-13)
    D2, D3 = D_(scheme)(img2), D_(scheme)(img3)
    nd2 = D2.shape[1]
    np.testing.assert_allclose(D3[3, :nd2], D2[0], atol=1e-14)
    assert np.all(D3[:, nd2:] == 0)
    np.testing.assert_allclose(DT_(scheme)(D3)[2], DT_(scheme)(D2)[0], atol=1e-13)


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("Nz,M", [(1, 2), (1, 8), (20, 3), (20, 4)])
def test_tv_D_DT_4D(scheme, Nz, M):
    """tests.py:304-361: TV, G, D, D^T D agree with the CPU implementation on 4-D data, reg_time = 1."""
    rs = np.random.RandomState(3)
    x = rs.rand(Nz, M, 40, 40)
    kw = dict(reg_time=1.0)
    tv, G = tv_(scheme)(x.copy(), **kw)
    tv_o, G_o = orc.tv(x.copy(), scheme, **kw)
    Dx, D_o = D_(scheme)(x, **kw), orc.D(x, scheme, **kw)
    assert float(tv) == pytest.approx(tv_o, rel=1e-13)
    assert float(opG.compute_L21_norm(Dx)) == pytest.approx(tv_o, rel=1e-13)
    np.testing.assert_allclose(G, G_o, atol=1e-11)
    np.testing.assert_allclose(Dx, D_o, atol=1e-14)
    np.testing.assert_allclose(DT_(scheme)(Dx, **kw), orc.D_T(D_o, scheme, **kw), atol=1e-12)


def test_central_nz2_and_small_volumes():
    rs = np.random.RandomState(9)
    x = rs.rand(2, 2, 6, 6)
    for kw in (dict(), dict(reg_time=0.5)):
        np.testing.assert_allclose(opG.D_central(x, **kw), orc.D(x, "central", **kw), atol=1e-14)
        tv, G = tvG.tv_central(x.copy(), **kw)
        tvo, Go = orc.tv(x.copy(), "central", **kw)
        assert float(tv) == pytest.approx(tvo, rel=1e-13)
        np.testing.assert_allclose(G, Go, atol=1e-12)
    # D_T_central on Nz = 3, N = 4: the reference's GPU path cannot, the CPU path is the oracle
    p = rs.randn(3, 3, 1, 4, 4)
    np.testing.assert_allclose(opG.D_T_central(p), orc.D_T(p, "central"), atol=1e-13)
    # 1x1 image, 2x2 image
    for N in (1, 2, 3):
        x = rs.rand(2, 2, N, N)
        for scheme in SCHEMES:
            if scheme == "central":
                continue
            tv, G = tv_(scheme)(x.copy(), reg_time=1.0)
            tvo, Go = orc.tv(x.copy(), scheme, reg_time=1.0)
            assert float(tv) == pytest.approx(tvo, rel=1e-12)
            np.testing.assert_allclose(G, Go, atol=1e-12)


def test_non_square_and_unaligned_paths():
    """Ni != Nj, and a device pointer that is not 16-byte aligned (forces the scalar kernels)."""
    rs = np.random.RandomState(13)
    x = rs.rand(3, 2, 5, 12).astype(np.float32)
    for scheme in SCHEMES:
        kw = dict(reg_time=0.5)
        np.testing.assert_allclose(D_(scheme)(x, **kw), orc.D(x, scheme, **kw), atol=1e-6)
    big = torch.rand(3 * 2 * 16 * 16 + 1, dtype=torch.float32, device="cuda")
    xa = big[:-1].view(3, 2, 16, 16)
    xu = big[1:].view(3, 2, 16, 16)          # misaligned by 4 bytes
    assert xu.data_ptr() % 16 != 0
    for scheme in SCHEMES:
        ref = orc.D(xu.cpu().numpy(), scheme, reg_time=0.5)
        np.testing.assert_allclose(D_(scheme)(xu, reg_time=0.5).cpu().numpy(), ref, atol=1e-6)
        tv_u, G_u = tv_(scheme)(xu.clone(), reg_time=0.5)
        tv_a, G_a = tv_(scheme)(xu.clone().contiguous(), reg_time=0.5)
        assert float(tv_u) == pytest.approx(float(tv_a), rel=1e-6)
        np.testing.assert_allclose(G_u, G_a, atol=1e-6)


@pytest.mark.parametrize("masked", [False, True], ids=["nomask", "mask_static"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("scheme", SCHEMES)
def test_tv_wide_rows_match_oracle(scheme, dtype, masked):
    """Rows wide enough that every warp of the TV sweeps covers 32 consecutive quads of one row (the lane exchange of
    sweep 2), 44 rows = five full strips of 8 and one that overhangs the image, both kernel variants of the time factor."""
    rs = np.random.RandomState(23)
    shape = (3, 3, 44, 1024)
    x = rs.rand(*shape).astype(dtype)
    kw = dict(reg_z_over_reg=0.7, reg_time=0.3)
    if masked:
        kw.update(mask_static=rs.rand(1, 1, shape[2], shape[3]) > 0.5, factor_reg_static=2.0)
    tv, G, n = tv_(scheme)(x.copy(), return_grad_norms=True, **kw)
    tv_o, G_o, n_o = orc.tv(x.astype(np.float64), scheme, return_grad_norms=True, **kw)
    f64 = dtype == np.float64
    assert float(tv) == pytest.approx(tv_o, rel=1e-13 if f64 else 1e-5)       # north-star tolerances for float32
    np.testing.assert_allclose(G, G_o, atol=1e-10 if f64 else 1e-5)
    np.testing.assert_allclose(n, n_o, atol=1e-13 if f64 else 1e-5)


@pytest.mark.parametrize("shape", [(3, 2, 101, 101), (2, 3, 37, 1001), (5, 1, 300, 130), (2, 2, 130, 258)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("scheme", SCHEMES)
def test_odd_and_multi_block_shapes(scheme, shape):
    """Row lengths not divisible by the vector width (scalar kernels), images spanning several row bands and
    column blocks; all operators in float64 against the oracle."""
    rs = np.random.RandomState(17)
    x = rs.rand(*shape)
    ms = rs.rand(1, 1, shape[2], shape[3]) > 0.5
    kw = dict(reg_z_over_reg=0.7, reg_time=0.3, mask_static=ms, factor_reg_static=2.0)
    D_o = orc.D(x, scheme, **kw)
    np.testing.assert_allclose(D_(scheme)(x, **kw), D_o, atol=1e-14)
    p = rs.randn(*D_o.shape)
    np.testing.assert_allclose(DT_(scheme)(p, **kw), orc.D_T(p, scheme, **kw), atol=1e-12)
    tv, G, n = tv_(scheme)(x.copy(), return_grad_norms=True, **kw)
    tv_o, G_o, n_o = orc.tv(x.copy(), scheme, return_grad_norms=True, **kw)
    assert float(tv) == pytest.approx(tv_o, rel=1e-13)
    np.testing.assert_allclose(G, G_o, atol=1e-10)
    np.testing.assert_allclose(n, n_o, atol=1e-13)
    s = pytv.CPSolver(x, lam=0.1, scheme=scheme, variant="rof", sigma=0.5, tau=0.07, **kw)
    s.step(2)
    xo, xb, y = x.copy(), x.copy(), np.zeros_like(D_o)
    for _ in range(2):
        xo, xb, y, e = orc.cp_rof_step(xo, xb, x, y, scheme, lam=0.1, sigma=0.5, tau=0.07, theta=1.0, **kw)
    np.testing.assert_allclose(s.x.cpu().numpy(), xo, atol=1e-12)
    assert s.energy() == pytest.approx(e, rel=1e-12)


@pytest.mark.parametrize("scheme", SCHEMES)
def test_time_weight_map(scheme, golden_small):
    """Extension (reference TODO, README.md:258): a (Nz,M,N,N) weight map of the time regularisation.  (1) the map
    where(mask, factor, 1) reproduces the reference's mask_static goldens; (2) a random map (with mask_static on top)
    matches the oracle for D, D_T, tv and CP, in float64 and float32; (3) adjointness."""
    case = [c for c in cases.small_cases() if c["key"] == "4x3x8/ztmask/" + scheme][0]
    x = cases.make_image(case)
    W = np.broadcast_to(np.where(cases.make_mask_static(case), case["fac"], 1.0), x.shape)
    kw = dict(reg_z_over_reg=case["rz"], reg_time=case["rt"], time_weight=W)
    np.testing.assert_allclose(D_(scheme)(x, **kw), golden_small[case["key"] + "/D"], atol=1e-13)
    tv, G = tv_(scheme)(x.copy(), **kw)
    assert float(tv) == pytest.approx(float(golden_small[case["key"] + "/tv"]), rel=1e-13)
    np.testing.assert_allclose(G, golden_small[case["key"] + "/G"], atol=1e-10)
    rs = np.random.RandomState(19)
    shape = (4, 5, 33, 68)
    x = rs.rand(*shape)
    W = rs.rand(*shape) * 3
    kw = dict(reg_z_over_reg=0.6, reg_time=0.4, mask_static=rs.rand(1, 1, 33, 68) > 0.5, factor_reg_static=2.0, time_weight=W)
    D_o = orc.D(x, scheme, **kw)
    Dx = D_(scheme)(x, **kw)
    np.testing.assert_allclose(Dx, D_o, atol=1e-14)
    p = rs.randn(*D_o.shape)
    DTp = DT_(scheme)(p, **kw)
    np.testing.assert_allclose(DTp, orc.D_T(p, scheme, **kw), atol=1e-12)
    assert np.sum(Dx * p) == pytest.approx(np.sum(x * DTp), rel=1e-12)
    tv, G, n = tv_(scheme)(x.copy(), return_grad_norms=True, **kw)
    tv_o, G_o, n_o = orc.tv(x.copy(), scheme, return_grad_norms=True, **kw)
    assert float(tv) == pytest.approx(tv_o, rel=1e-13)
    np.testing.assert_allclose(G, G_o, atol=1e-10)
    np.testing.assert_allclose(Dx.astype(np.float32), D_(scheme)(x.astype(np.float32), **kw), atol=1e-5)
    s = pytv.CPSolver(x, lam=0.1, scheme=scheme, variant="rof", sigma=0.5, tau=0.07, **kw)
    s.step(3)
    xo, xb, y = x.copy(), x.copy(), np.zeros_like(D_o)
    for _ in range(3):
        xo, xb, y, e = orc.cp_rof_step(xo, xb, x, y, scheme, lam=0.1, sigma=0.5, tau=0.07, theta=1.0, **kw)
    np.testing.assert_allclose(s.x.cpu().numpy(), xo, atol=1e-12)
    assert s.energy() == pytest.approx(e, rel=1e-12)


# ------------------------------------------------------------------ Chambolle-Pock
@pytest.mark.parametrize("shape", [(3, 2, 5, 8), (2, 3, 6, 4), (1, 1, 7, 12), (5, 1, 4, 8), (1, 4, 3, 4), (2, 2, 1, 4), (3, 3, 2, 8), (4, 2, 33, 260)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("scheme", SCHEMES)
def test_cp_single_iteration_all_shapes(scheme, shape):
    """One iteration from a random state (y non-zero also at the structurally-zero positions the adjoint must
    ignore) on small, degenerate and multi-CTA shapes; weights and mask_static on; float64 against the oracle."""
    lib = _lib.lib()
    rs = np.random.RandomState(31)
    Nz, M, Ni, Nj = shape
    x0 = rs.rand(*shape)
    ms = rs.rand(1, 1, Ni, Nj) > 0.5
    kw = dict(reg_z_over_reg=0.6, reg_time=0.4, mask_static=ms, factor_reg_static=3.0)
    Nd = orc.num_components(scheme, Nz, M, 0.6, 0.4)
    y = 0.2 * rs.randn(Nz, Nd, M, Ni, Nj)
    x = x0 + 0.1 * rs.randn(*shape)
    xbar = x + 0.01 * rs.randn(*shape)
    y_f = 0.1 * rs.randn(*shape)
    for variant in ("rof", "readme"):
        s = pytv.CPSolver(x0, lam=0.1, scheme=scheme, variant=variant, sigma=0.5, tau=0.07, theta=0.9, sigma_A=0.8, **kw)
        s.x.copy_(torch.as_tensor(x))
        s.y.copy_(torch.as_tensor(y))
        if variant == "rof":
            s.aux.copy_(torch.as_tensor(xbar))
            x_ref, aux_ref, y_ref, e_ref = orc.cp_rof_step(x.copy(), xbar.copy(), x0, y.copy(), scheme, lam=0.1, sigma=0.5, tau=0.07, theta=0.9, **kw)
        else:
            s.aux.copy_(torch.as_tensor(y_f))
            x_ref, aux_ref, y_ref, e_ref = orc.cp_readme_step(x.copy(), x0, y_f.copy(), y.copy(), scheme, lam=0.1, sigma_D=0.5, sigma_A=0.8, tau=0.07, **kw)
        s.step()
        np.testing.assert_allclose(s.y.cpu().numpy(), y_ref, atol=1e-12)
        np.testing.assert_allclose(s.x.cpu().numpy(), x_ref, atol=1e-12)
        np.testing.assert_allclose(s.aux.cpu().numpy(), aux_ref, atol=1e-12)
        assert s.energy() == pytest.approx(e_ref, rel=1e-12)


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_cp_small4d_golden(scheme, dtype, golden_kat):
    g = golden_kat["cp_small4d"][scheme]
    x0 = cases.cp_volume().astype(dtype)
    kw = dict(reg_z_over_reg=0.5, reg_time=2 ** -5, mask_static=cases.cp_mask_static(), factor_reg_static=4.0)
    rel = 1e-12 if dtype == np.float64 else 2e-5
    s = pytv.CPSolver(x0, lam=0.2, scheme=scheme, variant="readme", sigma=0.5, tau=1.0 / 17.0, sigma_A=1.0, **kw)
    losses = []
    for _ in range(10):
        s.step()
        losses.append(s.energy())
    np.testing.assert_allclose(losses, g["readme_losses"], rtol=rel)
    assert s.result().sum(dtype=np.float64) == pytest.approx(g["readme_sum_x"], rel=rel)
    assert float(s.y.abs().sum(dtype=torch.float64)) == pytest.approx(g["readme_sum_abs_y"], rel=10 * rel)
    s = pytv.CPSolver(x0, lam=0.2, scheme=scheme, variant="rof", sigma=0.5, tau=1.0 / 17.0, theta=1.0, **kw)
    energies = []
    for _ in range(10):
        s.step()
        energies.append(s.energy())
    np.testing.assert_allclose(energies, g["rof_energies"], rtol=rel)
    x = s.result()
    assert x.sum(dtype=np.float64) == pytest.approx(g["rof_sum_x"], rel=rel)
    assert float(s.aux.sum(dtype=torch.float64)) == pytest.approx(g["rof_sum_xbar"], rel=rel)
    assert x[1, 1, 3, 4] == pytest.approx(g["rof_x_probe"], rel=1e-11 if dtype == np.float64 else 1e-4)


@pytest.mark.parametrize("scheme", SCHEMES)
def test_half_precision_dual_storage(scheme):
    """y stored as normalised IEEE half.  Not a parity path - the bound checked here is the one the
    documentation states (max |dx| < 1e-3, rms < 1e-4 on [0,1] data after 200 iterations) - and it must equal a float32
    run whose dual field is rounded to half after every dual pass (i.e. storage is the ONLY difference)."""
    torch.manual_seed(0)
    N = 32
    blocks = (torch.arange(N) // 8) % 3
    x_true = ((blocks[:, None, None] + blocks[None, :, None] + blocks[None, None, :]) % 3 * 0.5).reshape(N, 1, N, N).float()
    x0 = (x_true + 0.1 * torch.randn(x_true.shape)).cuda()
    x0 = torch.cat([x0, x0.flip(0)], dim=1).contiguous()          # M = 2 so that the time axis is exercised too
    kw = dict(lam=0.1, scheme=scheme, variant="rof", reg_time=0.25)
    ref = pytv.CPSolver(x0, **kw)
    half = pytv.CPSolver(x0, dual_dtype=torch.float16, **kw)
    assert half.y.dtype == torch.float16 and half.y.shape == ref.y.shape
    emu = pytv.CPSolver(x0, **kw)
    for _ in range(200):
        ref.step(); half.step()
    for _ in range(3):                                            # storage-only difference, checked over 3 iterations
        emu._pass_A()
        emu.y.copy_(((emu.y / emu.lam).half().float()) * emu.lam)
        emu._pass_B()
    h3 = pytv.CPSolver(x0, dual_dtype=torch.float16, **kw)
    h3.step(3)
    # a rounding flip of one half-ulp of y moves x by tau*lam*2^-11 ~ 3.5e-6
    assert float((h3.x - emu.x).abs().max()) < 3e-5
    err = (half.x - ref.x).abs()
    assert float(err.max()) < 1e-3 and float((err ** 2).mean().sqrt()) < 1e-4, (float(err.max()), float((err ** 2).mean().sqrt()))
    assert half.energy() == pytest.approx(ref.energy(), rel=1e-3)
    with pytest.raises(ValueError):
        pytv.CPSolver(x0.double(), dual_dtype=torch.float16, **kw)


def test_cp_and_gd_loops_synthetic(golden_kat):
    """The README denoising loops (README.md:107-158) on the synthetic 64x64 image, 50 iterations."""
    x_true = cases.synthetic_image(64)
    noisy = x_true + 100 * np.random.RandomState(0).rand(*x_true.shape)
    g = golden_kat["gd_synthetic64"]
    x = noisy.copy()
    losses = []
    for _ in range(50):
        tv, G = tvG.tv_hybrid(x)
        x += -5e-3 * ((x - noisy) + 25.0 * G)
        losses.append(0.5 * np.sum(np.square(x - noisy)) + 25.0 * float(tv))
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-10)
    g = golden_kat["cp_readme_synthetic64"]
    s = pytv.CPSolver(noisy, lam=25.0, scheme="hybrid", variant="readme", sigma=0.5, sigma_A=1.0, tau=1.0 / 9.0)
    losses = []
    for _ in range(50):
        s.step()
        losses.append(s.energy())
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-11)
    assert s.result().sum() == pytest.approx(g["sum_x"], rel=1e-12)
    # the literal README loop through the operator API (numpy in, numpy out), 10 iterations
    x, y_f, y_tv = noisy.copy(), np.zeros_like(noisy), np.zeros((1, 4, 1, 64, 64))
    for it in range(10):
        y_f = (y_f + 1.0 * (x - noisy)) / 2.0
        D_x = opG.D_hybrid(x)
        pa = y_tv + 0.5 * D_x
        y_tv = pa / np.maximum(1.0, np.sqrt(np.sum(pa ** 2, axis=1)) / 25.0)
        x = x - (1 / 9.0) * y_f - (1 / 9.0) * opG.D_T_hybrid(y_tv)
        loss = 0.5 * np.sum(np.square(x - noisy)) + 25.0 * opG.compute_L21_norm(D_x)
        assert float(loss) == pytest.approx(g["losses"][it], rel=1e-11)


def test_gd_denoise_matches_readme_loop(golden_kat):
    """README.md:107-124 as a device-resident call: the first 50 losses of the reference loop (float64)."""
    x_true = cases.synthetic_image(64)
    noisy = x_true + 100 * np.random.RandomState(0).rand(*x_true.shape)
    g = golden_kat["gd_synthetic64"]
    x, losses = pytv.gd_denoise(noisy, 25.0, 50, 5e-3, scheme="hybrid", return_losses=True)
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-10)
    assert x.sum() == pytest.approx(g["sum_x"], rel=1e-11)


def test_cp_denoise_converges_and_reduces_energy():
    """BASELINE config 3 in miniature: piecewise-constant 3-D phantom + noise, hybrid, float32."""
    torch.manual_seed(0)
    N = 64
    blocks = (torch.arange(N) // 16) % 3
    x_true = (blocks[:, None, None] + blocks[None, :, None] + blocks[None, None, :]) % 3 * 0.5
    x_true = x_true.reshape(N, 1, N, N).float().cuda()
    x0 = x_true + 0.1 * torch.randn_like(x_true)
    s = pytv.CPSolver(x0, lam=0.1, scheme="hybrid", variant="rof")
    assert s.tau == pytest.approx(1 / 13.0)
    s.step(1)
    e_first = s.energy()
    s.step(199)
    e_last = s.energy()
    x = s.result(return_pytorch_tensor=True)
    assert e_last < e_first
    assert float(((x - x_true) ** 2).mean()) < 0.3 * float(((x0 - x_true) ** 2).mean())
    out = pytv.cp_denoise(x0, 0.1, 200, scheme="hybrid")
    assert torch.equal(out, x)        # deterministic


def test_cuda_graph_replay_equals_eager(golden_kat):
    """A captured iteration replayed 50 times reproduces the README CP losses on the synthetic 64x64 image."""
    x_true = cases.synthetic_image(64)
    noisy = x_true + 100 * np.random.RandomState(0).rand(*x_true.shape)
    g = golden_kat["cp_readme_synthetic64"]
    s = pytv.CPSolver(noisy, lam=25.0, scheme="hybrid", variant="readme", sigma=0.5, sigma_A=1.0, tau=1.0 / 9.0)
    s.capture_graph(iterations=1)
    losses = []
    for _ in range(50):
        s.step()
        losses.append(s.energy())
    assert s.iterations == 50
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-11)
    s2 = pytv.CPSolver(noisy, lam=25.0, scheme="hybrid", variant="readme", sigma=0.5, sigma_A=1.0, tau=1.0 / 9.0)
    s2.capture_graph(iterations=10)
    s2.step(50)
    assert s2.energy() == pytest.approx(g["losses"][-1], rel=1e-11)
    assert torch.equal(s2.x, s.x)


def test_pipelined_host_streaming_equals_synchronous():
    """step_host_async / wait (copy streams, double buffers) gives the same images and energies as step_host,
    with a different data term every step."""
    torch.manual_seed(11)
    shape = (6, 2, 64, 128)
    data = [torch.rand(shape).pin_memory() for _ in range(5)]
    a = pytv.CPSolver(data[0], lam=0.1, scheme="hybrid", variant="rof", reg_time=0.25)
    b = pytv.CPSolver(data[0], lam=0.1, scheme="hybrid", variant="rof", reg_time=0.25)
    outs_a = [torch.empty(shape).pin_memory() for _ in range(5)]
    outs_b = [torch.empty(shape).pin_memory() for _ in range(5)]
    e_a = [a.step_host(data[k], outs_a[k]) for k in range(5)]
    tickets, e_b = [], []
    for k in range(5):
        tickets.append(b.step_host_async(data[k], outs_b[k]))
        if k >= 1:
            e_b.append(b.wait(tickets[k - 1]))
    e_b.append(b.wait(tickets[-1]))
    assert e_a == e_b
    for k in range(5):
        assert torch.equal(outs_a[k], outs_b[k])


def test_primal_dual_gap_and_stopping_rule():
    """the gap P(x) - D(y) is non-negative, equals the oracle's value, shrinks, and stops `solve`."""
    rs = np.random.RandomState(5)
    x0 = cases.cp_volume() + 0.0
    kw = dict(reg_z_over_reg=0.5, reg_time=0.25)
    s = pytv.CPSolver(x0, lam=0.2, scheme="hybrid", variant="rof", **kw)
    s.step(5)
    p, d, g = s.gap()
    x, y = s.x.cpu().numpy(), s.y.cpu().numpy()
    p_o = 0.5 * np.sum((x - x0) ** 2) + 0.2 * orc.tv(x.copy(), "hybrid", **kw)[0]
    d_o = 0.5 * np.sum(x0 ** 2) - 0.5 * np.sum((x0 - orc.D_T(y, "hybrid", **kw)) ** 2)
    assert p == pytest.approx(p_o, rel=1e-12) and d == pytest.approx(d_o, rel=1e-12)
    assert g >= 0
    hist = []
    info = s.solve(max_iter=3000, tol=1e-6, check_every=50, callback=lambda sol, it, P, D: hist.append(P - D))
    assert info["converged"] and info["relative_gap"] <= 1e-6 and info["iterations"] < 3000
    assert hist[-1] < hist[0]
    # the solution is the TV prox of x0: TVProx with enough iterations lands on the same point
    prox = pytv.TVProx(x0, lam=0.2, scheme="hybrid", n_iter=info["iterations"] + 55, **kw)
    np.testing.assert_allclose(prox(x0).cpu().numpy(), s.x.cpu().numpy(), atol=1e-4)
    # data-term hook: a new x0 with a warm dual start converges in fewer iterations than a cold start
    x1 = x0 + 0.01 * rs.randn(*x0.shape)
    warm = s.set_data(x1).solve(max_iter=3000, tol=1e-6, check_every=10)
    cold = pytv.CPSolver(x1, lam=0.2, scheme="hybrid", variant="rof", **kw).solve(max_iter=3000, tol=1e-6, check_every=10)
    assert warm["converged"] and cold["converged"] and warm["iterations"] <= cold["iterations"]


def test_denoise_tv_chambolle_signature():
    """README.md:260 TODO: a skimage-compatible entry point.  Checked against a long oracle run of the same ROF
    problem (upwind = forward differences)."""
    img = cases.synthetic_image(64)[0, 0] / 255.0
    noisy = img + 0.1 * np.random.RandomState(1).randn(64, 64)
    out = pytv.denoise_tv_chambolle(noisy, weight=0.1, eps=1e-7, max_num_iter=2000)
    assert out.shape == noisy.shape and isinstance(out, np.ndarray)
    x0 = noisy.reshape(1, 1, 64, 64)
    x, xb, y = x0.copy(), x0.copy(), np.zeros((1, 2, 1, 64, 64))
    for _ in range(3000):
        x, xb, y, e = orc.cp_rof_step(x, xb, x0, y, "upwind", lam=0.1, sigma=0.5, tau=1.0 / 9.0)
    np.testing.assert_allclose(out, x[0, 0], atol=2e-3)
    assert np.mean((out - img) ** 2) < 0.5 * np.mean((noisy - img) ** 2)
    rgb = np.stack([noisy, noisy[::-1], noisy.T], axis=-1).astype(np.float32)
    out3 = pytv.denoise_tv_chambolle(rgb, weight=0.1, eps=0, max_num_iter=100, channel_axis=-1)
    assert out3.shape == rgb.shape and out3.dtype == np.float32
    np.testing.assert_allclose(out3[..., 0], pytv.denoise_tv_chambolle(rgb[..., 0], weight=0.1, eps=0, max_num_iter=100), atol=1e-5)
    vol = np.random.RandomState(2).rand(8, 16, 16)
    assert pytv.denoise_tv_chambolle(vol, weight=0.05, max_num_iter=20).shape == vol.shape


# ------------------------------------------------------------------ slabs through the real kernels
def _lib_call(fn, *a):
    _lib.check(fn(*a))


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
def test_slabs_with_halos_are_bitwise_equal_to_whole_volume(scheme, dtype):
    lib = _lib.lib()
    torch.manual_seed(5)
    Nz, M, N = 7, 3, 32
    x = torch.rand(Nz, M, N, N, dtype=dtype, device="cuda")
    ms = (torch.rand(N, N, device="cuda") > 0.5).to(torch.uint8)
    rz, rt, fac = 0.7, 0.3, 2.0
    did = _lib.F32 if dtype == torch.float32 else _lib.F64
    st = _dev.stream_ptr()

    def prob(a, b):
        return _lib.make_problem(scheme, did, (b - a, M, N, N), rz, rt, fac, ms.data_ptr(), a, Nz)

    whole = prob(0, Nz)
    Nd = lib.pytvb_num_components(ctypes.byref(whole))
    zf, zb = (4, 5) if scheme == "hybrid" else (2, 2)
    D_w = torch.empty(Nz, Nd, M, N, N, dtype=dtype, device="cuda")
    _lib_call(lib.pytvb_D, ctypes.byref(whole), _dev.ptr(x), _dev.ptr(D_w), None, None, st)
    p = torch.randn_like(D_w)
    DT_w = torch.empty_like(x)
    _lib_call(lib.pytvb_DT, ctypes.byref(whole), _dev.ptr(p), _dev.ptr(DT_w), None, None, st)
    G_w, n_w, tv_w = torch.empty_like(x), torch.empty_like(x), torch.zeros(1, dtype=torch.float64, device="cuda")
    wsr = _dev.reduce_workspace(whole, x.device)
    wst = torch.empty(lib.pytvb_tv_workspace_bytes(ctypes.byref(whole)), dtype=torch.uint8, device="cuda")
    _lib_call(lib.pytvb_tv, ctypes.byref(whole), _dev.ptr(x), _dev.ptr(G_w), _dev.ptr(n_w), _dev.ptr(tv_w), None, None, _dev.ptr(wsr), _dev.ptr(wst), st)
    tv_parts = 0.0
    for a, b in ((0, 2), (2, 3), (3, 7)):
        pb = prob(a, b)
        xs = x[a:b].contiguous()
        lo = x[a - 1].contiguous() if a > 0 else None
        hi = x[b].contiguous() if b < Nz else None
        D_s = torch.empty(b - a, Nd, M, N, N, dtype=dtype, device="cuda")
        _lib_call(lib.pytvb_D, ctypes.byref(pb), _dev.ptr(xs), _dev.ptr(D_s), _dev.ptr(lo), _dev.ptr(hi), st)
        assert torch.equal(D_s, D_w[a:b])
        plo = p[a - 1, zf].contiguous() if a > 0 else None
        phi = p[b, zb].contiguous() if b < Nz else None
        DT_s = torch.empty_like(xs)
        _lib_call(lib.pytvb_DT, ctypes.byref(pb), _dev.ptr(p[a:b].contiguous()), _dev.ptr(DT_s), _dev.ptr(plo), _dev.ptr(phi), st)
        assert torch.equal(DT_s, DT_w[a:b])
        lo2 = torch.full((2, M, N, N), float("nan"), dtype=dtype, device="cuda")
        hi2 = torch.full((2, M, N, N), float("nan"), dtype=dtype, device="cuda")
        for k in (1, 2):
            if a - k >= 0:
                lo2[2 - k] = x[a - k]
            if b + k - 1 < Nz:
                hi2[k - 1] = x[b + k - 1]
        G_s, n_s, tv_s = torch.empty_like(xs), torch.empty_like(xs), torch.zeros(1, dtype=torch.float64, device="cuda")
        _lib_call(lib.pytvb_tv, ctypes.byref(pb), _dev.ptr(xs), _dev.ptr(G_s), _dev.ptr(n_s), _dev.ptr(tv_s), _dev.ptr(lo2) if a > 0 else None,
                  _dev.ptr(hi2) if b < Nz else None, _dev.ptr(wsr), _dev.ptr(wst), st)
        assert torch.equal(G_s, G_w[a:b]) and torch.equal(n_s, n_w[a:b])
        tv_parts += float(tv_s[0])
    assert tv_parts == pytest.approx(float(tv_w[0]), rel=1e-12)


# ------------------------------------------------------------------ host-buffer C entry points
def test_host_entry_points():
    lib = _lib.lib()
    rs = np.random.RandomState(3)
    x = rs.rand(4, 3, 16, 16).astype(np.float32)
    ms = np.ascontiguousarray((rs.rand(16, 16) > 0.5).astype(np.uint8))
    pb = _lib.make_problem("hybrid", _lib.F32, x.shape, 0.5, 0.25, 4.0, ms.ctypes.data)
    G, norms, tv = np.empty_like(x), np.empty_like(x), ctypes.c_double(0)
    _lib.check(lib.pytvb_tv_host(ctypes.byref(pb), x.ctypes.data_as(ctypes.c_void_p), G.ctypes.data_as(ctypes.c_void_p),
                                 norms.ctypes.data_as(ctypes.c_void_p), ctypes.byref(tv)))
    tv_o, G_o = orc.tv(x.copy(), "hybrid", reg_z_over_reg=0.5, reg_time=0.25, mask_static=ms.reshape(1, 1, 16, 16).astype(bool), factor_reg_static=4.0)
    assert tv.value == pytest.approx(float(tv_o), rel=1e-5)
    _, G64 = orc.tv(x.astype(np.float64), "hybrid", reg_z_over_reg=0.5, reg_time=0.25, mask_static=ms.reshape(1, 1, 16, 16).astype(bool), factor_reg_static=4.0)
    _assert_G_float32(G, G64, x, "hybrid", dict(reg_z_over_reg=0.5, reg_time=0.25, mask_static=ms.reshape(1, 1, 16, 16).astype(bool), factor_reg_static=4.0))
    # streaming CP solver on host buffers == CPSolver on device buffers
    handle = ctypes.c_void_p()
    _lib.check(lib.pytvb_cp_create(ctypes.byref(pb), 0.2, 0.5, 1.0 / 17.0, 1.0, ctypes.byref(handle)))
    _lib.check(lib.pytvb_cp_reset_host(handle, x.ctypes.data_as(ctypes.c_void_p)))
    s = pytv.CPSolver(x, lam=0.2, scheme="hybrid", variant="rof", sigma=0.5, tau=1.0 / 17.0, reg_z_over_reg=0.5, reg_time=0.25,
                      mask_static=ms.reshape(1, 1, 16, 16).astype(bool), factor_reg_static=4.0)
    xo, e = np.empty_like(x), ctypes.c_double(0)
    for _ in range(5):
        _lib.check(lib.pytvb_cp_step_host(handle, x.ctypes.data_as(ctypes.c_void_p), xo.ctypes.data_as(ctypes.c_void_p), ctypes.byref(e)))
        s.step()
        assert e.value == pytest.approx(s.energy(), rel=1e-12)
    np.testing.assert_array_equal(xo, s.result())
    _lib.check(lib.pytvb_cp_destroy(handle))


# ------------------------------------------------------------------ launch-bound volumes: plans, graphs, BASELINE config 1
@pytest.mark.parametrize("scheme", SCHEMES)
def test_tv_plan_equals_the_drop_in_call(scheme):
    """TVPlan (pre-allocated buffers, CUDA-graph replay) returns exactly what tv_<scheme> returns, call after call, and the
    cached-plan path behind the numpy-in / numpy-out call equals the tensor path."""
    rs = np.random.RandomState(31)
    kw = dict(reg_z_over_reg=0.7, reg_time=0.3)
    plan = pytv.TVPlan(scheme, (5, 3, 33, 40), torch.float64, return_grad_norms=True, **kw)
    for k in range(3):
        x = rs.rand(5, 3, 33, 40)
        tv_p, G_p = plan(x)
        tv_t, G_t, n_t = tv_(scheme)(torch.as_tensor(x).cuda(), return_pytorch_tensor=True, return_grad_norms=True, **kw)
        assert float(tv_p) == float(tv_t)
        assert torch.equal(G_p, G_t) and torch.equal(plan.norms, n_t)
        tv_n, G_n, n_n = tv_(scheme)(x, return_grad_norms=True, **kw)          # numpy in -> cached plan
        assert float(tv_n) == float(tv_t) and np.array_equal(G_n, G_t.cpu().numpy()) and np.array_equal(n_n, n_t.cpu().numpy())
    x32 = rs.rand(5, 3, 33, 40).astype(np.float32)
    tv_n, G_n = tv_(scheme)(x32, **kw)
    tv_t, G_t = tv_(scheme)(torch.as_tensor(x32).cuda(), return_pytorch_tensor=True, **kw)
    assert G_n.dtype == np.float32 and float(tv_n) == float(tv_t) and np.array_equal(G_n, G_t.cpu().numpy())


def test_gd_denoise_graph_equals_eager(golden_kat):
    x_true = cases.synthetic_image(64)
    noisy = x_true + 100 * np.random.RandomState(0).rand(*x_true.shape)
    xa, la = pytv.gd_denoise(noisy, 25.0, 50, 5e-3, scheme="hybrid", return_losses=True, graph=True)
    xb, lb = pytv.gd_denoise(noisy, 25.0, 50, 5e-3, scheme="hybrid", return_losses=True, graph=False)
    assert np.array_equal(xa, xb) and np.array_equal(la, lb)
    np.testing.assert_allclose(la, golden_kat["gd_synthetic64"]["losses"], rtol=1e-10)
    with pytest.raises(TypeError):
        pytv.gd_denoise(noisy, 25.0, 2, 5e-3, no_such_weight=1.0)


def _cameraman():
    import os
    f = os.path.join(os.path.dirname(__file__), "..", "baseline", "_ref", "pytv", "media", "cameraman.npy")
    return np.load(f).astype(np.float64) if os.path.exists(f) else None


@pytest.mark.skipif(_cameraman() is None, reason="cameraman asset of the reference is not redistributed in this repo (travels in baseline/_ref)")
@pytest.mark.parametrize("N", [256, 512])
def test_config1_cameraman_300_iterations(N, golden_kat):
    """BASELINE config 1 as specified (README.md:107-158): the shipped 256 x 256 cameraman (and its 2x2 replication, 512 x 512),
    noise 100 rand (seed 0), lambda 25: 300 iterations of sub-gradient descent (step 5e-3) and of the README Chambolle-Pock loop
    on the GPU against the reference-generated goldens (256) / the oracle (512)."""
    cam = _cameraman()
    assert cam.sum() == golden_kat["cameraman_checksum"]["sum"]
    if N == 512:
        cam = np.kron(cam, np.ones((2, 2)))
    cam = cam.reshape(1, 1, N, N)
    np.random.seed(0)
    noisy = cam + 100 * np.random.rand(*cam.shape)
    x, losses = pytv.gd_denoise(noisy, 25.0, 300, 5e-3, scheme="hybrid", return_losses=True)
    s = pytv.CPSolver(noisy, lam=25.0, scheme="hybrid", variant="readme", tau=1 / 9.0)
    cp_losses = []
    for _ in range(300):
        s.step()
        cp_losses.append(s.energy())
    if N == 256:
        g = golden_kat["gd_cameraman"]
        assert losses[0] == pytest.approx(g["loss_first"], rel=1e-12)
        assert losses[-1] == pytest.approx(g["loss_last"], rel=1e-5)       # chaotic trajectory: see tests/test_oracle_golden.py
        c = golden_kat["cp_readme_cameraman"]
        assert cp_losses[0] == pytest.approx(c["loss_first"], rel=1e-12)
        assert cp_losses[-1] == pytest.approx(c["loss_last"], rel=1e-11)
        assert s.result().sum() == pytest.approx(c["sum_x"], rel=1e-12)
    else:
        xo = noisy.copy()
        for it in range(300):
            tv, G = orc.tv(xo, "hybrid")
            xo += -5e-3 * ((xo - noisy) + 25.0 * G)
            loss = 0.5 * np.sum(np.square(xo - noisy)) + 25.0 * tv
            if it < 40:
                assert losses[it] == pytest.approx(loss, rel=1e-11)
        assert losses[-1] == pytest.approx(loss, rel=1e-5)
        xr, y_f, y_tv = noisy.copy(), np.zeros_like(noisy), np.zeros((1, 4, 1, N, N))
        for it in range(300):
            xr, y_f, y_tv, loss = orc.cp_readme_step(xr, noisy, y_f, y_tv, "hybrid", lam=25.0)
        assert cp_losses[-1] == pytest.approx(loss, rel=1e-11)
    assert losses[-1] < 0.45 * losses[0] and cp_losses[-1] < losses[-1]        # README.md:126-162: CP converges faster than GD
