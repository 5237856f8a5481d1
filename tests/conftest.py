import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long CPU oracle runs, not part of the default CPU suite "
                                       "(LBM_RUN_SLOW=1 or -m slow runs them)")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("LBM_RUN_SLOW") == "1" or "slow" in (config.getoption("-m") or ""):
        return
    skip = pytest.mark.skip(reason="long CPU oracle run: set LBM_RUN_SLOW=1 (or -m slow)")
    for item in items:
        if "slow" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.load("base")
    return oracle_lib


@pytest.fixture(scope="session")
def lbm():
    """The product package; the CUDA library must already be built (make / build())."""
    import opencl_lattice_boltzmann_b200 as pkg
    pkg.cabi.load_library()
    return pkg
