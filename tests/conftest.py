import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: longer CPU test")
    config.addinivalue_line("markers", "fullsize: GPU parity at the real BASELINE.json sizes against committed golden "
                                       "vectors of the reference / oracle (tests/golden/cfg2_full_nfe32.npz, fullsize_*.npz)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _gpu_test_deadline(request):
    """Hard per-test deadline for GPU tests: a wedged kernel must not eat the GPU lease."""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    import faulthandler

    faulthandler.dump_traceback_later(900 if request.node.get_closest_marker("fullsize") else 240, exit=True)
    try:
        yield
    finally:
        faulthandler.cancel_dump_traceback_later()
