import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: minutes of CPU time; enabled with HYQUAS_SLOW=1")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("HYQUAS_SLOW") == "1":
        return
    skip = pytest.mark.skip(reason="set HYQUAS_SLOW=1 to run")
    for item in items:
        if "slow" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def gpu_runtime():
    """Binds the process to cuda:0 through the product's own init; fails loudly if the library is missing."""
    from hyquas_b200 import api
    api.init()
    return api
