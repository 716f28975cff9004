import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """Build libpytv_b200.so (and the host emulation harness) if a fresh checkout has not been built yet: nvcc
    cross-compiles for sm_100a without a GPU.  Source-newer-than-binary rebuilds are left to __graft_entry__.build()."""
    lib = os.path.join(ROOT, "pytv-4d_b200", "csrc", "libpytv_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden_small():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_small.npz"))


@pytest.fixture(scope="session")
def golden_kat():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden_kat.json")) as f:
        return json.load(f)
