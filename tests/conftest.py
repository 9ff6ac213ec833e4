"""pytest configuration: import paths, the ``gpu`` marker, shared fixtures.

``-m "not gpu"`` covers the oracle against its golden vectors / known answers, the host logic and
the C-ABI surface (library loads, exports every declared symbol; no compute calls).
``-m gpu`` runs the parity tests proper: every call goes through the C ABI on cuda:0.
Only tests (and smoke()/bench.py's CPU baseline) may import ``oracle``.
"""
import json
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (os.path.join(REPO, "robot-gym_b200"), REPO):
    if path not in sys.path:
        sys.path.insert(0, path)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def reference_constants():
    with open(os.path.join(GOLDEN, "reference_constants.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def rg_lib():
    """The C-ABI library, built in-tree if it is not there yet (nvcc cross-compiles without a GPU)."""
    from robot_gym import cuda as rg
    from robot_gym.cuda import build
    build.build()
    return rg.load()


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
