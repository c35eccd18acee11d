import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def bundled_model_paths():
    return sorted(glob.glob(os.path.join(GOLDEN, "models", "*.json")))


@pytest.fixture(scope="session")
def bundled_models():
    return bundled_model_paths()


@pytest.fixture(scope="session")
def native_libs():
    """Make sure the in-tree native libraries exist (they travel prebuilt to the GPU box)."""
    import aidadsp_lv2_b200 as ax
    if not os.path.exists(ax.lib_path("libaidax.so")):
        ax.build()
    return ax.LIB_DIR
