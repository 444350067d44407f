import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def hostsim():
    """ctypes handle of the TEST-ONLY host build of the element routines (tests/hostsim)."""
    import ctypes

    here = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(here, "_hostsim.so")
    src = os.path.join(here, "hostsim.cpp")
    deps = [src, os.path.join(ROOT, "flow_b200", "csrc", "fb_element.cuh"), os.path.join(ROOT, "flow_b200", "csrc", "fb_quadrature.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def gpu_ctx():
    from flow_b200 import _lib

    if not _lib.has_device():
        pytest.fail("flow_b200 found no CUDA device: the -m gpu tier must run on a GPU box (no CPU fallback exists)")
    return _lib.context()
