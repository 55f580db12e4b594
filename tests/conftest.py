import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def a2f_lib():
    """Build (no-op when up to date) and load liba2f_sm100.so."""
    import __graft_entry__ as g

    lib_path = os.path.join(ROOT, "audio2face-pytorch_b200", "liba2f_sm100.so")
    if not os.path.exists(lib_path):
        g.build()
    import a2f_b200

    return a2f_b200.lib.load()


@pytest.fixture(scope="session")
def dev():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
