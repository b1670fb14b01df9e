import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (oracle/libgrloracle.so) — the checker, built on demand."""
    from generalsreinforcementlearning_b200._abi import BoundLibrary

    path = os.path.join(ROOT, "oracle", "libgrloracle.so")
    src = os.path.join(ROOT, "oracle", "grl_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return BoundLibrary(path, "grlo_")


@pytest.fixture(scope="session")
def cuda_lib():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from generalsreinforcementlearning_b200 import load_library

    return load_library()
