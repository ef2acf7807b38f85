import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _build_product():
    """The in-tree native artefacts (CUDA library, C++ host library, rb) are built once per session — nvcc cross-compiles
    without a GPU — so that spawned worker processes and CLI tests find them."""
    from rustybam_b200 import build
    build.build_all()
    yield


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The CPU checker is built on demand (seconds); it is test infrastructure only."""
    so = os.path.join(ROOT, "oracle", "_build", "liborc.so")
    srcs = [os.path.join(ROOT, "oracle", f) for f in ("rb_oracle.cpp", "rb_oracle_capi.cpp", "rb_oracle.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    yield
