import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# the shared library is a build artefact (git-ignored): build it when the tests run on a clean tree
# (a machine without nvcc still collects and runs the oracle-only tests; the ones that load the library fail loudly there)
if not os.path.exists(os.path.join(ROOT, "bemstokes_b200", "libbemstokes_b200.so")):
    import importlib.util
    import shutil
    if shutil.which(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")) or shutil.which("nvcc"):
        _spec = importlib.util.spec_from_file_location("_bs_build", os.path.join(ROOT, "bemstokes_b200", "build.py"))
        _b = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_b)
        _b.build()
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
