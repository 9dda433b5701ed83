"""pytest configuration.

`-m "not gpu"` (CPU box): oracle vs golden vectors, host logic, C-ABI symbol checks, and the kernel sources
exercised through the HOST-SIMULATION build (oracle/hostsim.py, g++ -DFV3_HOSTSIM) — a test-only device that is
enabled here explicitly and can never be reached from the product API by accident.
`-m gpu` (B200): parity tests proper, through the CUDA C-ABI library.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

HAVE_GPU = torch.cuda.is_available()
if not HAVE_GPU:
    from oracle import hostsim as _hostsim

    _hostsim.install()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_ref = pytest.mark.skip(reason="/root/reference not present")
    have_ref = os.path.isdir("/root/reference/fv3core")
    for item in items:
        if "gpu" in item.keywords and not HAVE_GPU:
            item.add_marker(skip_gpu)
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def device():
    return "cuda" if HAVE_GPU else "cpu"
