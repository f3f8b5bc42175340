import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
MESHES = os.path.join(GOLDEN, "meshes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def goldens():
    import json
    with open(os.path.join(GOLDEN, "reference_goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def mesh_dir():
    return MESHES
