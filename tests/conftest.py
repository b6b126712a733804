"""Shared fixtures.  GPU tests are marked `gpu`; everything else runs on a CPU-only box.

The CPU checkers under oracle/ are used here only as checkers (see oracle/__init__.py).
"""
from __future__ import annotations

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


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def net():
    from stormphrax_b200 import net as N

    return N.synthetic(1234)


@pytest.fixture(scope="session")
def stress_net():
    from stormphrax_b200 import net as N

    return N.synthetic(99, stress=True)


@pytest.fixture(scope="session")
def c_oracle(net):
    from oracle.bind import COracle

    o = COracle()
    o.load_net(net.image)
    return o


@pytest.fixture(scope="session")
def reference(net):
    """The reference's own compiled code; absent on machines without a prebuilt oracle/_ref."""
    from oracle.bind import Reference

    if not Reference.available():
        pytest.skip("oracle/_ref not built / not runnable on this host")
    r = Reference()
    r.load_net(net.image)
    return r


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(GOLDEN, "playouts_seed42.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixtures missing: run tests/golden/make_golden.py where /root/reference exists")
    return np.load(path)


@pytest.fixture(scope="session")
def small_playouts():
    """~4k positions from the library's own playout generator (deterministic)."""
    from stormphrax_b200 import api

    return api.playouts(7, 50, 80, threads=4)


@pytest.fixture(scope="session")
def gpu_ctx(net):
    from stormphrax_b200 import api

    ctx = api.Nnue(net.image, 0)
    yield ctx
    ctx.close()
