"""pytest configuration: `gpu` marker + shared fixtures.

`-m "not gpu"` runs on the CPU build container (oracle vs golden vectors, host
logic, C-ABI symbol export, world_size-2 gloo sharding); `-m gpu` runs on a B200
and is the parity suite proper (every call goes through the C-ABI library).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def servo_kat():
    return load_golden("servo_kat.npz")


@pytest.fixture(scope="session")
def servo_chain():
    return load_golden("servo_chain.npz")


@pytest.fixture(scope="session")
def servo_edges():
    return load_golden("servo_edges.npz")


@pytest.fixture(scope="session")
def franka_golden():
    return load_golden("franka.npz")


@pytest.fixture(scope="session")
def franka_full():
    return load_golden("franka_full.npz")


@pytest.fixture(scope="session")
def pd_fragments():
    return load_golden("pd_fragments.npz")


def angle_diff_deg(a, b):
    """|a - b| on the circle, degrees."""
    d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) % 360.0
    return np.minimum(d, 360.0 - d)


def quat_diff(a, b):
    """max-abs difference of quaternions compared up to sign."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.minimum(np.abs(a - b).max(axis=-1), np.abs(a + b).max(axis=-1))
