import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "gpu_experimental: kernels that have never run on hardware yet (NOT part of -m gpu; "
                                       "run explicitly with -m gpu_experimental)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


HOST_HARNESS = os.environ.get("NPW_B200_HOST_HARNESS") == "1"


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if HOST_HARNESS:
        return torch.device("cpu")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(autouse=True)
def _host_harness(request, monkeypatch):
    """NPW_B200_HOST_HARNESS=1 (developer aid, never set by the driver): run the `-m gpu` tests' LOGIC on a machine
    without a GPU — CUDA streams/events replaced by inert stand-ins, the C-ABI by the NumPy double (tests/_fakecuda.py).
    It validates the tests and the host code they drive, not the kernels; results from such a run prove nothing about
    the GPU path and are never reported as GPU results."""
    if HOST_HARNESS and request.node.get_closest_marker("gpu") is not None:
        import _fakecuda
        _fakecuda.install(monkeypatch)
    yield


_counter = [0]


@pytest.fixture
def unique_key():
    def make(prefix="t"):
        _counter[0] += 1
        return "{0}_{1}_{2}".format(prefix, os.getpid(), _counter[0])
    return make
