import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests never silently pass on a CPU box: they are skipped unless selected with -m gpu,
    # and when selected they fail loudly if no device is visible.
    pass


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (test infrastructure) and make sure the product library exists."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True)
    lib = os.path.join(ROOT, "cuda_selection_criteria_b200", "libselb200.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "cuda_selection_criteria_b200", "csrc")], check=True)
    yield


@pytest.fixture(scope="session")
def gpu():
    import cuda_selection_criteria_b200 as S
    if S.lib().selb200_device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    return 0
