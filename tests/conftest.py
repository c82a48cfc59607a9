import os
import sys

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_lib():
    """Path of libigi_b200.so; built here when nvcc is available and it is stale/missing."""
    import shutil
    from isaacgyminsertion_b200 import build
    if shutil.which("nvcc"):
        build.build()
    assert os.path.exists(build.LIB), "libigi_b200.so missing and nvcc not available"
    return build.LIB


@pytest.fixture(scope="session", autouse=True)
def built_oracle():
    """The oracle's C parts (oracle/raster.c, oracle/fps.c) are test infrastructure: (re)built here with make when a
    compiler is available, so a fresh checkout does not depend on `__graft_entry__.build()` having run first."""
    import shutil
    import subprocess
    odir = os.path.join(ROOT, "oracle")
    if shutil.which("make") and (shutil.which("gcc") or shutil.which("cc")):
        subprocess.run(["make", "-C", odir], check=False, capture_output=True)   # a prebuilt copy is as good
    for so in ("liboracle_raster.so", "liboracle_fps.so"):
        assert os.path.exists(os.path.join(odir, so)), f"oracle/{so} missing and no compiler available"
