"""The per-thread CUDA code (pytv-4d_b200/csrc/strip_core.cuh, tile_core.cuh; and the retired generation-1 code kept under
tests/emul), executed on the host, against the oracle and the reference goldens.  This pins the index / boundary / halo / weight logic of the kernels
without a GPU; the -m gpu tests then check the same functions through the real launches."""
import numpy as np
import pytest

import cases
import emul_helper as em
from oracle import tv_oracle as orc

SCHEMES = cases.SCHEMES


@pytest.fixture(params=[(1, 8, 0, 0, 0), (2, 8, 0, 0, 0), (2, 4, 0, 0, 0), (3, 8, 0, 0, 1), (3, 8, 1, 4, 1), (3, 8, 0, 0, 2), (3, 8, 1, 4, 2)],
                ids=["gen1", "gen2", "gen2-r4", "tile", "tile-small", "tile2", "tile2-small"])
def gen(request):
    """The retired generation-1 quad code; the strip code (operators, CP passes, the two-sweep tv fallback with 8 and 4 rows per
    thread); and the strip code with tv through the single-sweep tile kernel - what the library runs - with the geometry the
    library chooses and with the smallest tiles and 4-plane z chunks (many CTAs, every seam exercised), in both forms of the tile
    kernel (two phases per plane / one phase)."""
    old = em.GEN
    em.GEN, rows, strips, Lz, form = request.param
    em.set_tv_rows(rows)
    em.set_tile(strips, Lz, form)
    yield em.GEN
    em.GEN = old
    em.set_tv_rows(8)
    em.set_tile(0, 0)


def _tol(dtype):
    return dict(rtol=0, atol=1e-12) if dtype == np.float64 else dict(rtol=0, atol=1e-5)


@pytest.mark.parametrize("scalar", [False, True], ids=["vec", "scalar"])
@pytest.mark.parametrize("case", [pytest.param(c, id=c["key"]) for c in cases.small_cases()])
def test_operators_match_reference_goldens(case, scalar, golden_small, gen):
    key, scheme = case["key"], case["scheme"]
    kw = cases.weight_kwargs(case)
    x = cases.make_image(case)
    gD = golden_small[key + "/D"]
    np.testing.assert_allclose(em.D(x, scheme, scalar=scalar, **kw), gD, rtol=0, atol=1e-13)
    p = cases.make_field(case, gD.shape)
    np.testing.assert_allclose(em.D_T(p, scheme, scalar=scalar, **kw), golden_small[key + "/DT"], rtol=0, atol=1e-11)
    tv, G, norms = em.tv(x, scheme, scalar=scalar, **kw)
    assert tv == pytest.approx(float(golden_small[key + "/tv"]), rel=1e-13)
    np.testing.assert_allclose(G, golden_small[key + "/G"], rtol=0, atol=1e-10)
    gn = golden_small[key + "/norms"]
    assert np.array_equal(np.isinf(norms), np.isinf(gn))
    np.testing.assert_allclose(norms[np.isfinite(gn)], gn[np.isfinite(gn)], rtol=0, atol=1e-13)


@pytest.mark.parametrize("case", [pytest.param(c, id=c["key"]) for c in cases.small_cases() if c["wname"] in ("default", "ztmask")])
def test_float32_within_north_star_tolerance(case, golden_small, gen):
    key, scheme = case["key"], case["scheme"]
    kw = cases.weight_kwargs(case)
    x = cases.make_image(case, np.float32)
    gD = golden_small[key + "/D"]
    np.testing.assert_allclose(em.D(x, scheme, **kw), gD, rtol=0, atol=1e-5)
    p = cases.make_field(case, gD.shape, np.float32)
    np.testing.assert_allclose(em.D_T(p, scheme, **kw), golden_small[key + "/DT"], rtol=0, atol=1e-5)
    tv, G, norms = em.tv(x, scheme, **kw)
    assert tv == pytest.approx(float(golden_small[key + "/tv"]), rel=1e-5)
    # against the float32 oracle the sub-gradient agrees to a few ulp even where norms are tiny
    _, G32 = orc.tv(x.copy(), scheme, **kw)
    np.testing.assert_allclose(G, G32, rtol=0, atol=5e-4)


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("split", [(2, 5), (1, 3, 4), (3,)], ids=["3slabs", "4slabs", "2slabs"])
def test_slabs_with_halos_reproduce_whole_volume(scheme, split, gen):
    """Nz-slab decomposition: every slab, given its halo planes, reproduces its part of the
    whole-volume result exactly."""
    rs = np.random.RandomState(21)
    Nz, M, N = 6, 3, 8
    x = rs.rand(Nz, M, N, N)
    x[..., :2, :3] = 0.5
    ms = rs.rand(1, 1, N, N) > 0.5
    kw = dict(reg_z_over_reg=0.7, reg_time=0.3, mask_static=ms, factor_reg_static=2.0)
    Dx = orc.D(x, scheme, **kw)
    p = rs.randn(*Dx.shape)
    DTp = orc.D_T(p, scheme, **kw)
    tv, G, norms = orc.tv(x.copy(), scheme, return_grad_norms=True, **kw)
    zf = {"upwind": 2, "downwind": 2, "central": 2, "hybrid": 4}[scheme]   # forward-type z slot
    zb = {"upwind": 2, "downwind": 2, "central": 2, "hybrid": 5}[scheme]   # backward-type z slot
    bounds = [0] + list(split) + [Nz]
    tv_sum = 0.0
    for a, b in zip(bounds[:-1], bounds[1:]):
        xs = np.ascontiguousarray(x[a:b])
        lo = np.ascontiguousarray(x[a - 1]) if a > 0 else None
        hi = np.ascontiguousarray(x[b]) if b < Nz else None
        np.testing.assert_allclose(em.D(xs, scheme, lo=lo, hi=hi, z_offset=a, Nz_global=Nz, **kw), Dx[a:b], atol=1e-14)
        plo = np.ascontiguousarray(p[a - 1, zf]) if a > 0 else None
        phi = np.ascontiguousarray(p[b, zb]) if b < Nz else None
        np.testing.assert_allclose(em.D_T(np.ascontiguousarray(p[a:b]), scheme, lo=plo, hi=phi, z_offset=a, Nz_global=Nz, **kw), DTp[a:b], atol=1e-13)
        # 2-plane halos for tv; planes outside the volume are never read, fill them with NaN to prove it
        lo2 = np.full((2, M, N, N), np.nan)
        hi2 = np.full((2, M, N, N), np.nan)
        for k in (1, 2):
            if a - k >= 0:
                lo2[2 - k] = x[a - k]
            if b + k - 1 < Nz:
                hi2[k - 1] = x[b + k - 1]
        tvs, Gs, ns = em.tv(xs, scheme, lo=lo2 if a > 0 else None, hi=hi2 if b < Nz else None, z_offset=a, Nz_global=Nz, **kw)
        np.testing.assert_allclose(Gs, G[a:b], atol=1e-12)
        np.testing.assert_allclose(ns, norms[a:b], atol=1e-13)
        tv_sum += tvs
    assert tv_sum == pytest.approx(tv, rel=1e-13)


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_cp_iterations_match_oracle(scheme, dtype, golden_kat, gen):
    g = golden_kat["cp_small4d"][scheme]
    x0 = cases.cp_volume().astype(dtype)
    kw = dict(reg_z_over_reg=0.5, reg_time=2 ** -5, mask_static=cases.cp_mask_static(), factor_reg_static=4.0)
    Nd = orc.num_components(scheme, 4, 3, 0.5, 2 ** -5)
    lam, sigma, tau = 0.2, 0.5, 1.0 / 17.0
    rel = 1e-12 if dtype == np.float64 else 2e-5
    # README form
    x, y_f, y = x0.copy(), np.zeros_like(x0), np.zeros((4, Nd, 3, 8, 8), dtype)
    losses = []
    for _ in range(10):
        l21 = em.cp_dual(x, y, scheme, lam, sigma, **kw)
        fid = em.cp_primal(y, x, y_f, x0, scheme, tau, 1.0, 1, **kw)
        losses.append(0.5 * fid + lam * l21)
    np.testing.assert_allclose(losses, g["readme_losses"], rtol=rel)
    assert x.sum(dtype=np.float64) == pytest.approx(g["readme_sum_x"], rel=rel)
    assert np.abs(y).sum(dtype=np.float64) == pytest.approx(g["readme_sum_abs_y"], rel=10 * rel)
    # ROF form
    x, xbar, y = x0.copy(), x0.copy(), np.zeros((4, Nd, 3, 8, 8), dtype)
    energies = []
    for _ in range(10):
        l21 = em.cp_dual(xbar, y, scheme, lam, sigma, **kw)
        fid = em.cp_primal(y, x, xbar, x0, scheme, tau, 1.0, 0, **kw)
        energies.append(0.5 * fid + lam * l21)
    np.testing.assert_allclose(energies, g["rof_energies"], rtol=rel)
    assert x.sum(dtype=np.float64) == pytest.approx(g["rof_sum_x"], rel=rel)
    assert xbar.sum(dtype=np.float64) == pytest.approx(g["rof_sum_xbar"], rel=rel)
    assert x[1, 1, 3, 4] == pytest.approx(g["rof_x_probe"], rel=1e-11 if dtype == np.float64 else 1e-4)


@pytest.mark.parametrize("shape", [(3, 2, 5, 8), (2, 3, 6, 4), (1, 1, 7, 12), (5, 1, 4, 8), (1, 4, 3, 4), (2, 2, 1, 4), (3, 3, 2, 8)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("scheme", SCHEMES)
def test_cp_single_iteration_all_shapes(scheme, shape, gen):
    """One iteration from a random state (non-zero y everywhere, also at the structurally-zero positions that
    the adjoint must ignore) on small and degenerate shapes; weights and mask_static on."""
    rs = np.random.RandomState(31)
    Nz, M, Ni, Nj = shape
    x0 = rs.rand(*shape)
    ms = rs.rand(1, 1, Ni, Nj) > 0.5
    kw = dict(reg_z_over_reg=0.6, reg_time=0.4, mask_static=ms, factor_reg_static=3.0)
    Nd = orc.num_components(scheme, Nz, M, 0.6, 0.4)
    y = 0.2 * rs.randn(Nz, Nd, M, Ni, Nj)
    x = x0 + 0.1 * rs.randn(*shape)
    xbar = x + 0.01 * rs.randn(*shape)
    for variant in (0, 1):
        if variant == 0:
            x_ref, aux_ref, y_ref, e_ref = orc.cp_rof_step(x.copy(), xbar.copy(), x0, y.copy(), scheme, lam=0.1, sigma=0.5, tau=0.07, theta=0.9, **kw)
            yy, xx, aux = y.copy(), x.copy(), xbar.copy()
            l21 = em.cp_dual(xbar, yy, scheme, 0.1, 0.5, **kw)
            fid = em.cp_primal(yy, xx, aux, x0, scheme, 0.07, 0.9, 0, **kw)
        else:
            y_f = 0.1 * rs.randn(*shape)
            x_ref, aux_ref, y_ref, e_ref = orc.cp_readme_step(x.copy(), x0, y_f.copy(), y.copy(), scheme, lam=0.1, sigma_D=0.5, sigma_A=0.8, tau=0.07, **kw)
            yy, xx, aux = y.copy(), x.copy(), y_f.copy()
            l21 = em.cp_dual(x, yy, scheme, 0.1, 0.5, **kw)
            fid = em.cp_primal(yy, xx, aux, x0, scheme, 0.07, 0.8, 1, **kw)
        np.testing.assert_allclose(yy, y_ref, atol=1e-13)
        np.testing.assert_allclose(xx, x_ref, atol=1e-13)
        np.testing.assert_allclose(aux, aux_ref, atol=1e-13)
        assert 0.5 * fid + 0.1 * l21 == pytest.approx(e_ref, rel=1e-12)


@pytest.mark.parametrize("scalar", [False, True], ids=["vec", "scalar"])
@pytest.mark.parametrize("scheme", SCHEMES)
def test_half_precision_dual_storage(scheme, scalar):
    """y stored as y/lam in IEEE half: identical to the float32 oracle iteration with the normalised dual rounded to
    half after every dual pass (storage is the only difference)."""
    rs = np.random.RandomState(8)
    shape = (3, 2, 6, 8)
    x0 = rs.rand(*shape).astype(np.float32)
    kw = dict(reg_z_over_reg=0.6, reg_time=0.4)
    Nd = orc.num_components(scheme, 3, 2, 0.6, 0.4)
    lam, sigma, tau, theta = 0.15, 0.5, 0.07, 1.0
    x, xbar, yh = x0.copy(), x0.copy(), np.zeros((3, Nd, 2, 6, 8), np.float16)
    xo, xbo, yo = x0.copy(), x0.copy(), np.zeros((3, Nd, 2, 6, 8), np.float32)
    for _ in range(5):
        l21, fid = em.cp_step_f16y(xbar, yh, x, x0, scheme, lam, sigma, tau, theta, scalar=scalar, **kw)
        Dxb = orc.D(xbo, scheme, **kw)
        yn = orc.project_l2_ball(yo + np.float32(sigma) * Dxb, lam)
        yo = (yn / np.float32(lam)).astype(np.float16).astype(np.float32) * np.float32(lam)
        xn = (xo - np.float32(tau) * orc.D_T(yo, scheme, **kw) + np.float32(tau) * x0) / np.float32(1 + tau)
        xbo = xn + np.float32(theta) * (xn - xo)
        xo = xn
        assert l21 == pytest.approx(float(orc.l21(Dxb)), rel=1e-5)
    np.testing.assert_allclose(x, xo, atol=3e-6)
    np.testing.assert_allclose(yh.astype(np.float32) * lam, yo, atol=2e-4 * lam + 1e-7)


@pytest.mark.parametrize("scheme", SCHEMES)
def test_cp_slabs(scheme, gen):
    """One CP iteration computed slab by slab with halos equals the whole-volume iteration."""
    rs = np.random.RandomState(4)
    Nz, M, N = 5, 2, 8
    x0 = rs.rand(Nz, M, N, N)
    kw = dict(reg_z_over_reg=0.5, reg_time=0.25)
    Nd = orc.num_components(scheme, Nz, M, 0.5, 0.25)
    y = 0.05 * rs.randn(Nz, Nd, M, N, N)
    x = x0 + 0.1 * rs.randn(*x0.shape)
    xbar = x + 0.01 * rs.randn(*x0.shape)
    x_ref, xbar_ref, y_ref, e_ref = orc.cp_rof_step(x.copy(), xbar.copy(), x0, y.copy(), scheme, lam=0.1, sigma=0.5, tau=0.07, theta=0.9, **kw)
    zf = 4 if scheme == "hybrid" else 2
    zb = 5 if scheme == "hybrid" else 2
    cut = 2
    y_new = y.copy()
    l21 = 0.0
    for a, b in ((0, cut), (cut, Nz)):
        ys = np.ascontiguousarray(y_new[a:b])
        lo = np.ascontiguousarray(xbar[a - 1]) if a > 0 else None
        hi = np.ascontiguousarray(xbar[b]) if b < Nz else None
        l21 += em.cp_dual(np.ascontiguousarray(xbar[a:b]), ys, scheme, 0.1, 0.5, lo=lo, hi=hi, z_offset=a, Nz_global=Nz, **kw)
        y_new[a:b] = ys
    np.testing.assert_allclose(y_new, y_ref, atol=1e-13)
    x_new, xbar_new = x.copy(), xbar.copy()
    fid = 0.0
    for a, b in ((0, cut), (cut, Nz)):
        xs, xbs = np.ascontiguousarray(x_new[a:b]), np.ascontiguousarray(xbar_new[a:b])
        plo = np.ascontiguousarray(y_new[a - 1, zf]) if a > 0 else None
        phi = np.ascontiguousarray(y_new[b, zb]) if b < Nz else None
        fid += em.cp_primal(np.ascontiguousarray(y_new[a:b]), xs, xbs, np.ascontiguousarray(x0[a:b]), scheme, 0.07, 0.9, 0, lo=plo, hi=phi,
                            z_offset=a, Nz_global=Nz, **kw)
        x_new[a:b], xbar_new[a:b] = xs, xbs
    np.testing.assert_allclose(x_new, x_ref, atol=1e-13)
    np.testing.assert_allclose(xbar_new, xbar_ref, atol=1e-13)
    assert 0.5 * fid + 0.1 * l21 == pytest.approx(e_ref, rel=1e-12)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cp_peer_halo_push(scheme, variant, dtype):
    """Multi-GPU schedule without an exchange step: every "rank" (a slab in this process) stores its boundary planes
    straight into its neighbours' halo buffers from inside the passes (the MIR kernels); three iterations are bitwise
    equal to the whole-volume strip code.  The slabs run one after the other, a barrier between passes is implied."""
    rs = np.random.RandomState(11)
    Nz, M, N = 6, 2, 8
    bounds = [(0, 1), (1, 4), (4, 6)]          # a one-plane slab too: plane 0 is also its last plane
    kw = dict(reg_z_over_reg=0.5, reg_time=0.25)
    Nd = orc.num_components(scheme, Nz, M, 0.5, 0.25)
    x0 = rs.rand(Nz, M, N, N).astype(dtype)
    x = x0.copy()
    aux = x0.copy() if variant == 0 else np.zeros_like(x0)
    y = np.zeros((Nz, Nd, M, N, N), dtype=dtype)
    lam, sigma, tau, c2 = 0.1, 0.5, 0.07, (0.9 if variant == 0 else 1.0)
    # whole volume
    xw, auxw, yw = x.copy(), aux.copy(), y.copy()
    saved, em.GEN = em.GEN, 2
    try:
        for _ in range(3):
            em.cp_dual(auxw if variant == 0 else xw, yw, scheme, lam, sigma, **kw)
            em.cp_primal(yw, xw, auxw, x0, scheme, tau, c2, variant, **kw)
    finally:
        em.GEN = saved
    # slabs with four halo planes each: [img_lo, img_hi, fld_lo, fld_hi]
    nr = len(bounds)
    X = [np.ascontiguousarray(x[a:b]) for a, b in bounds]
    A = [np.ascontiguousarray(aux[a:b]) for a, b in bounds]
    Y = [np.ascontiguousarray(y[a:b]) for a, b in bounds]
    X0 = [np.ascontiguousarray(x0[a:b]) for a, b in bounds]
    H = [np.full((4, M, N, N), np.nan, dtype=dtype) for _ in bounds]
    need_img_lo, need_img_hi = scheme != "upwind", scheme != "downwind"
    need_fld_lo, need_fld_hi = scheme != "downwind", scheme != "upwind"
    U = A if variant == 0 else X
    for r in range(nr):                        # the one start-up exchange of the image halos
        if r > 0:
            H[r][0] = U[r - 1][-1]
        if r < nr - 1:
            H[r][1] = U[r + 1][0]
    for _ in range(3):
        for r, (a, b) in enumerate(bounds):
            mp = H[r - 1][3] if (r > 0 and need_fld_hi) else None        # my backward-type z slot -> previous rank's fld_hi
            mn = H[r + 1][2] if (r < nr - 1 and need_fld_lo) else None   # my forward-type z slot -> next rank's fld_lo
            em.cp_dual_mirror(U[r], Y[r], scheme, lam, sigma, H[r][0] if (r > 0 and need_img_lo) else None,
                              H[r][1] if (r < nr - 1 and need_img_hi) else None, mp, mn, a, Nz, **kw)
        for r, (a, b) in enumerate(bounds):
            mp = H[r - 1][1] if (r > 0 and need_img_hi) else None        # my first plane -> previous rank's img_hi
            mn = H[r + 1][0] if (r < nr - 1 and need_img_lo) else None   # my last plane -> next rank's img_lo
            em.cp_primal_mirror(Y[r], X[r], A[r], X0[r], scheme, tau, c2, variant, H[r][2] if (r > 0 and need_fld_lo) else None,
                                H[r][3] if (r < nr - 1 and need_fld_hi) else None, mp, mn, a, Nz, **kw)
    np.testing.assert_array_equal(np.concatenate(X), xw)
    np.testing.assert_array_equal(np.concatenate(A), auxw)
    np.testing.assert_array_equal(np.concatenate(Y), yw)


def test_central_nz2_intent(gen):
    rs = np.random.RandomState(9)
    x = rs.rand(2, 2, 6, 6)
    for kw in (dict(), dict(reg_time=0.5)):
        np.testing.assert_allclose(em.D(x, "central", **kw), orc.D(x, "central", **kw), atol=1e-14)
        p = rs.randn(*orc.D(x, "central", **kw).shape)
        np.testing.assert_allclose(em.D_T(p, "central", **kw), orc.D_T(p, "central", **kw), atol=1e-13)
        tv, G, _ = em.tv(x, "central", **kw)
        tvo, Go = orc.tv(x.copy(), "central", **kw)
        assert tv == pytest.approx(tvo, rel=1e-13)
        np.testing.assert_allclose(G, Go, atol=1e-12)


@pytest.mark.parametrize("Ni", [1, 2, 3, 7, 8, 9, 15, 17])
@pytest.mark.parametrize("scheme", SCHEMES)
def test_tv_image_heights_around_the_strip_height(scheme, Ni, gen):
    """The row-marching TV sweeps with images shorter than, equal to and just past a multiple of the strip height."""
    rs = np.random.RandomState(31 + Ni)
    x = rs.rand(3, 2, Ni, 8)
    ms = rs.rand(1, 1, Ni, 8) > 0.5
    for kw in (dict(reg_time=0.5), dict(reg_z_over_reg=0.3, reg_time=0.7, mask_static=ms, factor_reg_static=3.0)):
        tv, G, n = em.tv(x, scheme, **kw)
        tvo, Go, no = orc.tv(x.copy(), scheme, return_grad_norms=True, **kw)
        assert tv == pytest.approx(tvo, rel=1e-13)
        np.testing.assert_allclose(G, Go, atol=1e-12)
        np.testing.assert_allclose(n, no, atol=1e-13)


def test_non_square_images(gen):
    """The kernels take Ni and Nj separately (reference: square only, README.md:259)."""
    rs = np.random.RandomState(13)
    x = rs.rand(3, 2, 5, 12)
    for scheme in SCHEMES:
        kw = dict(reg_time=0.5)
        Dx = orc.D(x, scheme, **kw)
        np.testing.assert_allclose(em.D(x, scheme, **kw), Dx, atol=1e-14)
        p = rs.randn(*Dx.shape)
        np.testing.assert_allclose(em.D_T(p, scheme, **kw), orc.D_T(p, scheme, **kw), atol=1e-13)
        tv, G, _ = em.tv(x, scheme, **kw)
        tvo, Go = orc.tv(x.copy(), scheme, **kw)
        assert tv == pytest.approx(tvo, rel=1e-13)
        np.testing.assert_allclose(G, Go, atol=1e-12)


# ------------------------------------------------------------------ extension: per-voxel weight map of the time axis
@pytest.mark.parametrize("scalar", [False, True], ids=["vec", "scalar"])
@pytest.mark.parametrize("shape", [(3, 4, 6, 8), (1, 2, 5, 4), (4, 3, 7, 12), (2, 5, 3, 8)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("scheme", SCHEMES)
def test_time_weight_map(scheme, shape, scalar):
    """Strip kernels with a (Nz,M,N,N) weight map of the time regularisation (+ mask_static on top) against the oracle:
    D, exact adjoint, TV / sub-gradient, and one CP iteration of both forms."""
    old = em.GEN
    em.GEN = 2
    try:
        rs = np.random.RandomState(19)
        Nz, M, Ni, Nj = shape
        x = rs.rand(*shape)
        W = rs.rand(*shape) * 3
        ms = rs.rand(1, 1, Ni, Nj) > 0.5
        kw = dict(reg_z_over_reg=0.6, reg_time=0.4, mask_static=ms, factor_reg_static=2.0, time_weight=W)
        D_o = orc.D(x, scheme, **kw)
        np.testing.assert_allclose(em.D(x, scheme, scalar=scalar, **kw), D_o, atol=1e-14)
        p = rs.randn(*D_o.shape)
        np.testing.assert_allclose(em.D_T(p, scheme, scalar=scalar, **kw), orc.D_T(p, scheme, **kw), atol=1e-13)
        tv_o, G_o, n_o = orc.tv(x.copy(), scheme, return_grad_norms=True, **kw)
        # two-sweep fallback, tile kernel (both forms), tile kernel with small tiles
        for g_, strips, Lz, form in ((2, 0, 0, 0), (3, 0, 0, 1), (3, 1, 4, 1), (3, 0, 0, 2), (3, 1, 4, 2)):
            em.GEN = g_
            em.set_tile(strips, Lz, form)
            tv, G, n = em.tv(x, scheme, scalar=scalar, **kw)
            assert tv == pytest.approx(tv_o, rel=1e-13)
            np.testing.assert_allclose(G, G_o, atol=1e-12)
            np.testing.assert_allclose(n, n_o, atol=1e-13)
        # slabs with z halos (tile kernel): the weight map travels with one halo plane per side (pytvb_problem.time_scale_lo / _hi)
        for form in ((1, 2) if Nz >= 3 and not (scheme == "central" and M == 2) else ()):
            em.GEN = 3
            em.set_tile(0, 0, form)
            tv_sum = 0.0
            for a, b in ((0, 1), (1, Nz)):
                lo2 = np.full((2, M, Ni, Nj), np.nan)
                hi2 = np.full((2, M, Ni, Nj), np.nan)
                for k in (1, 2):
                    if a - k >= 0:
                        lo2[2 - k] = x[a - k]
                    if b + k - 1 < Nz:
                        hi2[k - 1] = x[b + k - 1]
                kws = dict(kw, time_weight=W[a:b])
                tvs, Gs, ns = em.tv(np.ascontiguousarray(x[a:b]), scheme, scalar=scalar, lo=lo2 if a > 0 else None, hi=hi2 if b < Nz else None, z_offset=a,
                                    Nz_global=Nz, time_scale_halos=(W[a - 1] if a > 0 else None, W[b] if b < Nz else None), **kws)
                np.testing.assert_allclose(Gs, G_o[a:b], atol=1e-12)
                np.testing.assert_allclose(ns, n_o[a:b], atol=1e-13)
                tv_sum += tvs
            assert tv_sum == pytest.approx(tv_o, rel=1e-13)
        em.GEN = 2
        em.set_tile(0, 0)
        Nd = D_o.shape[1]
        y = 0.2 * rs.randn(Nz, Nd, M, Ni, Nj)
        xx = x + 0.1 * rs.randn(*shape)
        xbar = xx + 0.01 * rs.randn(*shape)
        x_ref, xb_ref, y_ref, e_ref = orc.cp_rof_step(xx.copy(), xbar.copy(), x, y.copy(), scheme, lam=0.1, sigma=0.5, tau=0.07, theta=0.9, **kw)
        yy, x2, aux = y.copy(), xx.copy(), xbar.copy()
        l21 = em.cp_dual(xbar, yy, scheme, 0.1, 0.5, scalar=scalar, **kw)
        fid = em.cp_primal(yy, x2, aux, x, scheme, 0.07, 0.9, 0, scalar=scalar, **kw)
        np.testing.assert_allclose(yy, y_ref, atol=1e-13)
        np.testing.assert_allclose(x2, x_ref, atol=1e-13)
        np.testing.assert_allclose(aux, xb_ref, atol=1e-13)
        assert 0.5 * fid + 0.1 * l21 == pytest.approx(e_ref, rel=1e-12)
    finally:
        em.GEN = old


# ------------------------------------------------------------------ tile-kernel geometry (host logic of kernels_tile.cuh)
def test_tile_kernel_form_and_geometry_for_the_baseline_configs():
    """Which form of the single-sweep kernel the library picks, and with which tile, for the BASELINE configurations: the one-phase
    form (4 x slots + 2 w windows) for up to four coupled frames, the two-phase form where the fourth slot would cost tile rows;
    the shared-memory footprint stays inside the opt-in limit; z chunks fill the 148 SMs in whole waves."""
    g = em.tile_geometry((128, 4, 1024, 1024), np.float32, reg_time=2 ** -5)                  # C4 slab
    assert (g["form"], g["FC"], g["strips"], g["TI"], g["TJ"], g["nthreads"]) == (2, 4, 4, 14, 120, 512)
    assert g["smem"] <= g["smem_limit"] <= 227 * 1024 and g["smem"] > 200 * 1024
    assert g["nblocks"] == 74 * 9 * g["nzc"] and g["nblocks"] % 148 == 0 and g["Lz"] * g["nzc"] >= 128
    g = em.tile_geometry((64, 8, 2048, 2048), np.float32, reg_time=2 ** -5)                   # C5 slab: eight coupled frames
    assert (g["form"], g["FC"], g["strips"], g["TI"]) == (1, 8, 2, 6)
    assert g["smem"] <= g["smem_limit"]
    g = em.tile_geometry((64, 8, 2048, 2048), np.float32, reg_time=2 ** -5, mask_static=np.ones((2048, 2048), bool), fac=4.0)
    assert (g["form"], g["strips"]) == (1, 2) and g["smem"] <= g["smem_limit"]                 # the mask-factor tile still fits
    g = em.tile_geometry((512, 1, 512, 512), np.float32)                                      # C3: no time axis, 16 strips of one frame
    assert (g["form"], g["FC"], g["strips"], g["TI"]) == (2, 1, 16, 62) and g["smem"] <= g["smem_limit"]
    g = em.tile_geometry((20, 4, 100, 100), np.float64, reg_time=2 ** -5)                     # C2 (README volume, float64)
    assert g["form"] == 2 and g["TJ"] == 60 and g["Lz"] <= 4 and g["nblocks"] <= 148          # short z chunks: one wave of CTAs
    g = em.tile_geometry((1, 1, 256, 256), np.float32)                                        # C1: one plane
    assert g["form"] in (1, 2) and g["nzc"] == 1 and g["smem"] <= g["smem_limit"]
    g = em.tile_geometry((4, 32, 64, 64), np.float32, reg_time=1.0)                           # more coupled frames than the CTA has warps
    assert g["form"] == 0                                                                     # -> two-sweep fallback
