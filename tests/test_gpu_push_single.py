"""The `_p2p` entry points (pytvb_cp_dual_p2p / pytvb_cp_primal_p2p: passes that store their boundary planes into the
neighbours' halo planes) on ONE GPU: the "ranks" are z-slabs of one volume on the same device, their halo planes ordinary
device buffers, and the passes of the slabs run one after the other on the stream (tests/p2p_schedule.py).  The result
must equal the whole-volume iteration bit for bit.  The same schedule across two GPUs, with the planes in symmetric
memory, is tests/test_gpu_multigpu.py."""
import pytest
import torch

import p2p_schedule
from pytv_b200 import cp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", ["rof", "readme"])
@pytest.mark.parametrize("scheme", ["upwind", "downwind", "central", "hybrid"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
def test_push_schedule_on_one_device(scheme, variant, dtype):
    slabs, whole = p2p_schedule.run(cp.CudaOps(), torch.device("cuda", 0), scheme, variant, dtype)
    torch.cuda.synchronize()
    for a, b in zip(slabs, whole):
        assert torch.equal(a, b)


def test_push_schedule_wide_planes():
    """Planes wide enough for the 128-bit path with several column blocks and row bands; uneven slabs."""
    slabs, whole = p2p_schedule.run(cp.CudaOps(), torch.device("cuda", 0), "hybrid", "rof", torch.float32, shape=(7, 3, 70, 520),
                                    bounds=((0, 2), (2, 3), (3, 7)), iterations=2)
    torch.cuda.synchronize()
    for a, b in zip(slabs, whole):
        assert torch.equal(a, b)
