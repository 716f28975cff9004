"""world_size-2 (and 3) runs of the slab-sharded Chambolle-Pock solver over the gloo backend on CPU.
The per-pass arithmetic is the CUDA per-quad code executed on the host (tests/emul); what is under test is
the N>1 host logic of pytv_b200.cp: slab placement, which planes go to which neighbour, halo buffers, the
energy all-reduce.  The result must equal the single-domain oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, scheme, variant, Nz, out_dir, overlap):
    os.environ["PYTVB_OVERLAP"] = overlap
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import emul_helper as em
    import pytv_b200
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(17)
        M, N = 2, 8
        x0 = rs.rand(Nz, M, N, N)
        ms = rs.rand(1, 1, N, N) > 0.5
        off, cnt = pytv_b200.partition_z(Nz, world)[rank]
        solver = pytv_b200.CPSolver(x0[off:off + cnt], lam=0.15, scheme=scheme, variant=variant, reg_z_over_reg=0.5, reg_time=0.25,
                                    mask_static=ms, factor_reg_static=2.0, distributed=True, ops=em.EmulOps())
        assert solver.z_offset == off and solver.Nz_global == Nz
        energies = []
        for _ in range(4):
            solver.step()
            h = solver.energy_async()
            energies.append(solver.energy())
            assert solver.energy_result(h) == energies[-1]
        np.save(os.path.join(out_dir, "x_%d.npy" % rank), solver.result())
        np.save(os.path.join(out_dir, "y_%d.npy" % rank), solver.y.numpy())
        if rank == 0:
            np.save(os.path.join(out_dir, "energies.npy"), np.array(energies))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", ["0", "1"], ids=["blocking", "overlap"])
@pytest.mark.parametrize("variant", ["rof", "readme"])
@pytest.mark.parametrize("scheme,world,Nz", [("hybrid", 2, 5), ("upwind", 2, 4), ("downwind", 2, 4), ("central", 3, 5), ("hybrid", 3, 4)])
def test_sharded_cp_equals_single_domain(tmp_path, scheme, world, Nz, variant, overlap):
    from oracle import tv_oracle as orc
    if variant == "readme" and not (scheme == "hybrid" and world == 2):
        pytest.skip("README form is exercised with the hybrid scheme (its reference loop, README.md:145-157)")
    if overlap == "1" and scheme in ("upwind", "downwind"):
        pytest.skip("the overlapped schedule is exercised with the two-sided schemes")
    port = _free_port()
    mp.spawn(_worker, args=(world, port, scheme, variant, Nz, str(tmp_path), overlap), nprocs=world, join=True)
    x = np.concatenate([np.load(tmp_path / ("x_%d.npy" % r)) for r in range(world)], axis=0)
    y = np.concatenate([np.load(tmp_path / ("y_%d.npy" % r)) for r in range(world)], axis=0)
    energies = np.load(tmp_path / "energies.npy")
    rs = np.random.RandomState(17)
    M, N = 2, 8
    x0 = rs.rand(Nz, M, N, N)
    ms = rs.rand(1, 1, N, N) > 0.5
    kw = dict(reg_z_over_reg=0.5, reg_time=0.25, mask_static=ms, factor_reg_static=2.0)
    Nd = orc.num_components(scheme, Nz, M, 0.5, 0.25)
    L2 = (1.0 if scheme == "central" else 4.0) * (2 + 0.5 + 0.25 * 2.0)
    tau = 1.0 / (L2 + 1.0)
    ref_e = []
    if variant == "rof":
        xr, xb, yr = x0.copy(), x0.copy(), np.zeros((Nz, Nd, M, N, N))
        for _ in range(4):
            xr, xb, yr, e = orc.cp_rof_step(xr, xb, x0, yr, scheme, lam=0.15, sigma=0.5, tau=tau, theta=1.0, **kw)
            ref_e.append(e)
    else:
        xr, yf, yr = x0.copy(), np.zeros_like(x0), np.zeros((Nz, Nd, M, N, N))
        for _ in range(4):
            xr, yf, yr, e = orc.cp_readme_step(xr, x0, yf, yr, scheme, lam=0.15, sigma_D=0.5, sigma_A=1.0, tau=tau, **kw)
            ref_e.append(e)
    np.testing.assert_allclose(x, xr, atol=1e-13)
    np.testing.assert_allclose(y, yr, atol=1e-13)
    np.testing.assert_allclose(energies, ref_e, rtol=1e-12)


def _sharded_worker(rank, world, port, scheme, Nz, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import emul_helper as em
    import pytv_b200
    from pytv_b200.sharded import ShardedTV
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(23)
        M, N = 2, 8
        x = rs.rand(Nz, M, N, N)
        x[:, :, :2, :3] = 0.5
        ms = rs.rand(1, 1, N, N) > 0.5
        mask = rs.rand(N, N) > 0.2
        off, cnt = pytv_b200.partition_z(Nz, world)[rank]
        sh = ShardedTV(scheme, reg_z_over_reg=0.7, reg_time=0.3, mask_static=ms, factor_reg_static=2.0, ops=em.EmulSlabOps())
        xs = torch.as_tensor(x[off:off + cnt].copy())
        Ds = sh.D(xs)
        p = rs.randn(Nz, Ds.shape[1], M, N, N)
        DTs = sh.D_T(torch.as_tensor(p[off:off + cnt].copy()))
        l21 = sh.l21(Ds)
        tv, G, norms = sh.tv(xs.clone(), return_grad_norms=True)
        tvm, Gm = sh.tv(xs, mask=torch.as_tensor(mask))
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), D=Ds.numpy(), DT=DTs.numpy(), G=G.numpy(), norms=norms.numpy(), Gm=Gm.numpy(), xm=xs.numpy(),
                 scal=np.array([l21, tv, tvm]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("scheme,world,Nz", [("hybrid", 2, 5), ("central", 3, 7), ("upwind", 2, 4), ("downwind", 2, 4)])
def test_sharded_operators_and_tv_equal_whole_volume(tmp_path, scheme, world, Nz):
    """pytv_b200.sharded.ShardedTV over gloo: D, D_T, L21 and tv (value, sub-gradient, norms, in-place mask) of the slabs
    equal the whole-volume oracle."""
    from oracle import tv_oracle as orc
    mp.spawn(_sharded_worker, args=(world, _free_port(), scheme, Nz, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / ("r%d.npz" % r)) for r in range(world)]
    cat = lambda k: np.concatenate([p[k] for p in parts], axis=0)
    rs = np.random.RandomState(23)
    M, N = 2, 8
    x = rs.rand(Nz, M, N, N)
    x[:, :, :2, :3] = 0.5
    ms = rs.rand(1, 1, N, N) > 0.5
    mask = rs.rand(N, N) > 0.2
    kw = dict(reg_z_over_reg=0.7, reg_time=0.3, mask_static=ms, factor_reg_static=2.0)
    D_o = orc.D(x, scheme, **kw)
    p = rs.randn(*D_o.shape)
    np.testing.assert_allclose(cat("D"), D_o, atol=1e-14)
    np.testing.assert_allclose(cat("DT"), orc.D_T(p, scheme, **kw), atol=1e-13)
    tv_o, G_o, n_o = orc.tv(x.copy(), scheme, return_grad_norms=True, **kw)
    for part in parts:
        assert part["scal"][0] == pytest.approx(float(orc.l21(D_o)), rel=1e-13)
        assert part["scal"][1] == pytest.approx(tv_o, rel=1e-13)
    np.testing.assert_allclose(cat("G"), G_o, atol=1e-12)
    np.testing.assert_allclose(cat("norms"), n_o, atol=1e-13)
    xm = x.copy()
    tvm_o, Gm_o = orc.tv(xm, scheme, mask=np.broadcast_to(mask, x.shape), **kw)
    np.testing.assert_array_equal(cat("xm"), xm)
    np.testing.assert_allclose(cat("Gm"), Gm_o, atol=1e-12)
    assert parts[0]["scal"][2] == pytest.approx(tvm_o, rel=1e-13)


def _weighted_worker(rank, world, port, scheme, Nz, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import emul_helper as em
    import pytv_b200
    from pytv_b200.sharded import ShardedTV
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(29)
        M, N = 3, 8
        x = rs.rand(Nz, M, N, N)
        W = rs.rand(Nz, M, N, N) * 3
        ms = rs.rand(1, 1, N, N) > 0.5
        off, cnt = pytv_b200.partition_z(Nz, world)[rank]
        sh = ShardedTV(scheme, reg_z_over_reg=0.7, reg_time=0.3, mask_static=ms, factor_reg_static=2.0, ops=em.EmulSlabOps())
        xs = torch.as_tensor(x[off:off + cnt].copy())
        Ws = torch.as_tensor(W[off:off + cnt].copy())
        Ds = sh.D(xs, time_weight=Ws)
        p = rs.randn(Nz, Ds.shape[1], M, N, N)
        DTs = sh.D_T(torch.as_tensor(p[off:off + cnt].copy()), time_weight=Ws)
        tv, G, norms = sh.tv(xs.clone(), return_grad_norms=True, time_weight=Ws)
        tv0, G0 = sh.tv(xs.clone())
        np.savez(os.path.join(out_dir, "w%d.npz" % rank), D=Ds.numpy(), DT=DTs.numpy(), G=G.numpy(), norms=norms.numpy(), G0=G0.numpy(), scal=np.array([tv, tv0]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("scheme,world,Nz", [("hybrid", 2, 14), ("central", 2, 13), ("upwind", 3, 20)])
def test_sharded_interior_split_and_weight_map(tmp_path, scheme, world, Nz):
    """Slabs of >= 6 planes: the interior planes are computed before the halos arrive, the two boundary planes on each side
    after (ShardedTV's overlapped schedule); with a (Nz, M, N, N) weight map of the time regularisation, whose boundary
    planes travel with the image planes (pytvb_problem.time_scale_lo / _hi)."""
    from oracle import tv_oracle as orc
    mp.spawn(_weighted_worker, args=(world, _free_port(), scheme, Nz, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / ("w%d.npz" % r)) for r in range(world)]
    cat = lambda k: np.concatenate([p[k] for p in parts], axis=0)
    rs = np.random.RandomState(29)
    M, N = 3, 8
    x = rs.rand(Nz, M, N, N)
    W = rs.rand(Nz, M, N, N) * 3
    ms = rs.rand(1, 1, N, N) > 0.5
    kw = dict(reg_z_over_reg=0.7, reg_time=0.3, mask_static=ms, factor_reg_static=2.0)
    D_o = orc.D(x, scheme, time_weight=W, **kw)
    p = rs.randn(*D_o.shape)
    np.testing.assert_allclose(cat("D"), D_o, atol=1e-14)
    np.testing.assert_allclose(cat("DT"), orc.D_T(p, scheme, time_weight=W, **kw), atol=1e-13)
    tv_o, G_o, n_o = orc.tv(x.copy(), scheme, return_grad_norms=True, time_weight=W, **kw)
    np.testing.assert_allclose(cat("G"), G_o, atol=1e-12)
    np.testing.assert_allclose(cat("norms"), n_o, atol=1e-13)
    tv0_o, G0_o = orc.tv(x.copy(), scheme, **kw)
    np.testing.assert_allclose(cat("G0"), G0_o, atol=1e-12)
    for part in parts:
        assert part["scal"][0] == pytest.approx(tv_o, rel=1e-13) and part["scal"][1] == pytest.approx(tv0_o, rel=1e-13)
