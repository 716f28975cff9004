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


def _worker(rank, world, port, scheme, out_dir, overlap):
    os.environ["PYTVB_OVERLAP"] = overlap
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
        s = pytv_b200.CPSolver(x0[off:off + cnt].cuda(), lam=0.1, scheme=scheme, variant="rof", reg_time=0.25, distributed=True)
        energies = []
        for _ in range(5):
            s.step()
            energies.append(s.energy())
        np.save(os.path.join(out_dir, "x_%d.npy" % rank), s.result())
        if rank == 0:
            np.save(os.path.join(out_dir, "e.npy"), np.array(energies))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", ["0", "1"], ids=["blocking", "overlap"])
@pytest.mark.parametrize("scheme", ["hybrid", "central", "upwind"])
def test_two_gpu_sharded_cp_equals_single_gpu(tmp_path, scheme, overlap):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import pytv_b200
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), scheme, str(tmp_path), overlap), nprocs=world, join=True)
    x = np.concatenate([np.load(tmp_path / ("x_%d.npy" % r)) for r in range(world)], axis=0)
    g = torch.Generator().manual_seed(3)
    x0 = torch.rand(12, 3, 64, 64, generator=g)
    s = pytv_b200.CPSolver(x0.cuda(), lam=0.1, scheme=scheme, variant="rof", reg_time=0.25)
    energies = []
    for _ in range(5):
        s.step()
        energies.append(s.energy())
    np.testing.assert_array_equal(x, s.result())
    np.testing.assert_allclose(np.load(tmp_path / "e.npy"), energies, rtol=1e-12)
