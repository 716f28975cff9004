"""Host-side logic of pytv_b200.cp that needs neither a GPU nor a process group."""
import numpy as np
import pytest

import emul_helper as em
import pytv_b200
from pytv_b200 import cp


def test_comm_argument_is_validated():
    x0 = np.random.RandomState(0).rand(2, 2, 4, 4)
    with pytest.raises(ValueError):
        pytv_b200.CPSolver(x0, lam=0.1, ops=em.EmulOps(), comm="mpi")
    # a single-process solver has no halos: every transport name is accepted and nothing is set up
    for comm in ("auto", "nccl", "p2p"):
        s = pytv_b200.CPSolver(x0, lam=0.1, ops=em.EmulOps(), comm=comm)
        assert s.halo is None and s._peer is None
        s.step(1)


def test_peer_halo_slot_addresses():
    """PeerHalos.peer: slot k of rank r is r's mapped base address + k planes; no neighbour -> NULL."""
    class Handle:
        buffer_ptrs = [0x1000_0000, 0x2000_0000, 0x3000_0000]

    p = object.__new__(cp.PeerHalos)
    p.hdl, p.plane_bytes = Handle(), 4 * 2 * 8 * 8
    assert p.peer(None, cp.PeerHalos.FLD_HI) is None
    assert p.peer(0, cp.PeerHalos.IMG_LO) == 0x1000_0000
    assert p.peer(2, cp.PeerHalos.FLD_HI) == 0x3000_0000 + 3 * p.plane_bytes
    assert (cp.PeerHalos.IMG_LO, cp.PeerHalos.IMG_HI, cp.PeerHalos.FLD_LO, cp.PeerHalos.FLD_HI) == (0, 1, 2, 3)


def test_partition_z_covers_the_volume():
    for Nz in (1, 5, 8, 1024):
        for world in (1, 2, 3, 8):
            if world > Nz:
                continue
            parts = pytv_b200.partition_z(Nz, world)
            assert parts[0][0] == 0 and sum(c for _, c in parts) == Nz
            assert all(parts[k][0] + parts[k][1] == parts[k + 1][0] for k in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


@pytest.mark.parametrize("variant", ["rof", "readme"])
@pytest.mark.parametrize("scheme", ["upwind", "downwind", "central", "hybrid"])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_push_schedule_through_the_executor_interface(scheme, variant, dtype):
    """tests/p2p_schedule.py (the harness the single-GPU test of the `_p2p` entry points uses) with the host-emulated
    executor: slabs pushing their boundary planes into each other's halo planes equal the whole volume bit for bit."""
    import torch
    import p2p_schedule
    slabs, whole = p2p_schedule.run(em.EmulOps(), torch.device("cpu"), scheme, variant, getattr(torch, dtype))
    for a, b in zip(slabs, whole):
        assert torch.equal(a, b)


def test_push_schedule_uneven_slabs_wide_planes():
    import torch
    import p2p_schedule
    slabs, whole = p2p_schedule.run(em.EmulOps(), torch.device("cpu"), "hybrid", "rof", torch.float32, shape=(7, 3, 70, 520),
                                    bounds=((0, 2), (2, 3), (3, 7)), iterations=2)
    for a, b in zip(slabs, whole):
        assert torch.equal(a, b)
