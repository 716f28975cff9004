"""Two-rank NCCL run of the sharded Chambolle-Pock solver against the single-GPU run (needs >= 2 GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, scheme, out_dir, mode, variant="rof", weighted=False, sync="auto"):
    os.environ["PYTVB_OVERLAP"] = "1" if mode == "overlap" else "0"
    os.environ["PYTVB_P2P_SYNC"] = sync
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import pytv_b200
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g = torch.Generator().manual_seed(3)
        x0 = torch.rand(12, 3, 64, 64, generator=g)
        off, cnt = pytv_b200.partition_z(12, world)[rank]
        tw = (torch.rand(12, 3, 64, 64, generator=g) * 3)[off:off + cnt].cuda() if weighted else None
        s = pytv_b200.CPSolver(x0[off:off + cnt].cuda(), lam=0.1, scheme=scheme, reg_time=0.25, distributed=True,
                                comm=mode if mode in ("p2p", "auto") else "nccl", variant=variant, time_weight=tw)
        if mode != "auto":        # auto: peer memory where the box offers it, NCCL otherwise - the result is the same
            assert (s._peer is not None) == (mode == "p2p")
        energies = []
        for _ in range(5):
            s.step()
            energies.append(s.energy())
        if variant == "rof":      # the diagnostics reuse the halo buffers between iterations: iterate on afterwards
            energies.extend(s.gap())
            s.step(2)
        np.save(os.path.join(out_dir, "x_%d.npy" % rank), s.result())
        if rank == 0:
            np.save(os.path.join(out_dir, "e.npy"), np.array(energies))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["blocking", "overlap", "p2p", "auto"])
@pytest.mark.parametrize("scheme", ["hybrid", "central", "upwind", "downwind"])
def test_two_gpu_sharded_cp_equals_single_gpu(tmp_path, scheme, mode):
    """mode: NCCL send/recv before each pass | the same hidden behind the interior planes | no exchange at all, the
    kernels store the boundary planes into the neighbour's halo buffers over NVLink (peer memory)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import pytv_b200
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), scheme, str(tmp_path), mode), nprocs=world, join=True)
    x = np.concatenate([np.load(tmp_path / ("x_%d.npy" % r)) for r in range(world)], axis=0)
    g = torch.Generator().manual_seed(3)
    x0 = torch.rand(12, 3, 64, 64, generator=g)
    s = pytv_b200.CPSolver(x0.cuda(), lam=0.1, scheme=scheme, variant="rof", reg_time=0.25)
    energies = []
    for _ in range(5):
        s.step()
        energies.append(s.energy())
    energies.extend(s.gap())
    s.step(2)
    np.testing.assert_array_equal(x, s.result())
    e = np.load(tmp_path / "e.npy")
    np.testing.assert_allclose(e[:5], energies[:5], rtol=1e-12)
    np.testing.assert_allclose(e[5:7], energies[5:7], rtol=1e-9)            # primal, dual energy (other summation grouping)
    assert abs(e[7] - energies[7]) <= 1e-9 * abs(energies[5])


def _sharded_worker(rank, world, port, scheme, out_dir, comm="nccl", Nz=10, weighted=False):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import pytv_b200
    from pytv_b200.sharded import ShardedTV
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g = torch.Generator().manual_seed(5)
        x = torch.rand(Nz, 3, 64, 64, generator=g)
        ms = torch.rand(1, 1, 64, 64, generator=g) > 0.5
        W = torch.rand(Nz, 3, 64, 64, generator=g) * 3
        off, cnt = pytv_b200.partition_z(Nz, world)[rank]
        sh = ShardedTV(scheme, reg_z_over_reg=0.7, reg_time=0.3, mask_static=ms, factor_reg_static=2.0, comm=comm)
        xs = x[off:off + cnt].cuda()
        tw = W[off:off + cnt].cuda() if weighted else None
        for rep in range(3):      # repeated calls: the peer buffers alternate, a rank may run one call ahead of its neighbour
            Ds = sh.D(xs, time_weight=tw)
            p = torch.randn(Nz, Ds.shape[1], 3, 64, 64, generator=torch.Generator().manual_seed(6))
            DTs = sh.D_T(p[off:off + cnt].cuda(), time_weight=tw)
            l21 = sh.l21(Ds)
            tv, G, norms = sh.tv(xs, return_grad_norms=True, time_weight=tw)
        assert sh.transport == ("p2p" if comm == "p2p" else sh.transport)
        np.savez(os.path.join(out_dir, "s%d.npz" % rank), D=Ds.cpu().numpy(), DT=DTs.cpu().numpy(), G=G.cpu().numpy(), norms=norms.cpu().numpy(),
                 scal=np.array([l21, tv]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("comm,Nz,weighted", [("nccl", 10, False), ("p2p", 10, False), ("p2p", 17, True), ("nccl", 17, True)],
                         ids=["nccl", "p2p", "p2p-interior-weighted", "nccl-interior-weighted"])
@pytest.mark.parametrize("scheme", ["hybrid", "central", "downwind"])
def test_two_gpu_sharded_operators_equal_single_gpu(tmp_path, scheme, comm, Nz, weighted):
    """ShardedTV on two GPUs - halo planes by NCCL send/recv or stored straight into the neighbour's peer-mapped buffers,
    slabs of 5 planes (whole-slab calls) and of 8-9 planes (interior first, boundary planes after the halos landed), with and
    without a weight map of the time regularisation: bitwise the single-GPU drop-in results on the whole volume."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import pytv_b200 as pytv
    mp.spawn(_sharded_worker, args=(2, _free_port(), scheme, str(tmp_path), comm, Nz, weighted), nprocs=2, join=True)
    parts = [np.load(tmp_path / ("s%d.npz" % r)) for r in range(2)]
    cat = lambda k: np.concatenate([q[k] for q in parts], axis=0)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(Nz, 3, 64, 64, generator=g)
    ms = torch.rand(1, 1, 64, 64, generator=g) > 0.5
    W = torch.rand(Nz, 3, 64, 64, generator=g) * 3
    kw = dict(reg_z_over_reg=0.7, reg_time=0.3, mask_static=ms, factor_reg_static=2.0)
    if weighted:
        kw["time_weight"] = W.cuda()
    D1 = getattr(pytv.tv_operators_GPU, "D_" + scheme)(x.cuda(), **kw)
    p = torch.randn(Nz, D1.shape[1], 3, 64, 64, generator=torch.Generator().manual_seed(6))
    DT1 = getattr(pytv.tv_operators_GPU, "D_T_" + scheme)(p.cuda(), **kw)
    tv1, G1, n1 = getattr(pytv.tv_GPU, "tv_" + scheme)(x.cuda(), return_pytorch_tensor=True, return_grad_norms=True, **kw)
    np.testing.assert_array_equal(cat("D"), D1.cpu().numpy())
    np.testing.assert_array_equal(cat("DT"), DT1.cpu().numpy())
    np.testing.assert_array_equal(cat("G"), G1.cpu().numpy())
    np.testing.assert_array_equal(cat("norms"), n1.cpu().numpy())
    assert parts[0]["scal"][1] == pytest.approx(float(tv1), rel=1e-6)
    assert parts[0]["scal"][0] == pytest.approx(float(pytv.tv_operators_GPU.compute_L21_norm(D1)), rel=1e-6)


def test_two_gpu_peer_halo_push_readme_variant(tmp_path):
    """The README form of the iteration (x, y_f, y_tv) through the peer-memory halo push: pass B pushes x, not xbar."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import pytv_b200
    mp.spawn(_worker, args=(2, _free_port(), "hybrid", str(tmp_path), "p2p", "readme"), nprocs=2, join=True)
    x = np.concatenate([np.load(tmp_path / ("x_%d.npy" % r)) for r in range(2)], axis=0)
    g = torch.Generator().manual_seed(3)
    x0 = torch.rand(12, 3, 64, 64, generator=g)
    s = pytv_b200.CPSolver(x0.cuda(), lam=0.1, scheme="hybrid", variant="readme", reg_time=0.25)
    energies = []
    for _ in range(5):
        s.step()
        energies.append(s.energy())
    np.testing.assert_array_equal(x, s.result())
    np.testing.assert_allclose(np.load(tmp_path / "e.npy"), energies, rtol=1e-12)


@pytest.mark.parametrize("sync", ["auto", "barrier"])
def test_two_gpu_peer_halo_push_with_weight_map(tmp_path, sync):
    """Peer-memory halo push together with a (Nz, M, N, N) weight map of the time regularisation (the combination round 1
    refused), with the neighbour-only handshake and with the group-wide barrier as the fence between passes."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import pytv_b200
    mp.spawn(_worker, args=(2, _free_port(), "hybrid", str(tmp_path), "p2p", "rof", True, sync), nprocs=2, join=True)
    x = np.concatenate([np.load(tmp_path / ("x_%d.npy" % r)) for r in range(2)], axis=0)
    g = torch.Generator().manual_seed(3)
    x0 = torch.rand(12, 3, 64, 64, generator=g)
    tw = torch.rand(12, 3, 64, 64, generator=g) * 3
    s = pytv_b200.CPSolver(x0.cuda(), lam=0.1, scheme="hybrid", variant="rof", reg_time=0.25, time_weight=tw.cuda())
    energies = []
    for _ in range(5):
        s.step()
        energies.append(s.energy())
    energies.extend(s.gap())
    s.step(2)
    np.testing.assert_array_equal(x, s.result())
    np.testing.assert_allclose(np.load(tmp_path / "e.npy")[:5], energies[:5], rtol=1e-12)
