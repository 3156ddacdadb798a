import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


@pytest.fixture(autouse=True, scope="session")
def _single_thread_torch_cpu():
    """The oracle's float coordinates are pinned with one CPU thread (see oracle/gen_golden.py: torch
    mixes SLEEF and scalar libm per thread chunk, so (u, v) of a few tail elements depend on it)."""
    import torch
    torch.set_num_threads(1)
    yield


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cuda_lib():
    """Builds (if needed) and loads the C-ABI library; GPU tests fail loudly without it."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from delora_b200 import _lib
    return _lib.lib()
