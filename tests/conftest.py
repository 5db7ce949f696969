"""Test plumbing.  `-m "not gpu"` runs here on CPU: the oracle against the golden vectors,
host-side logic, and that the C-ABI library loads and exports what include/planet_gpu.h
declares.  `-m gpu` are the parity tests proper: CUDA path (through the C-ABI) vs oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by the reference's own code (oracle/gen_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "planet_golden.npz"))


@pytest.fixture(scope="session")
def port():
    from oracle.bindings import PortOracle
    return PortOracle()


@pytest.fixture(scope="session")
def ref():
    """The reference compiled as a library -- only where oracle/_ref was built."""
    from oracle.bindings import RefOracle
    if not RefOracle.available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return RefOracle()


@pytest.fixture(scope="session")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device: the product has no CPU path")
    import planet_b200 as pb
    pb.init(0)
    return pb


def as_bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def quads_from_bytes(b):
    from oracle.bindings import QUAD_DTYPE
    return np.ascontiguousarray(b).reshape(-1).view(QUAD_DTYPE)
