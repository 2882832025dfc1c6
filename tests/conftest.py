import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def harness_lib():
    """Host build of the device math (tests/host_harness), compiled with g++ on first use."""
    import ctypes
    import subprocess
    hdir = os.path.join(ROOT, "tests", "host_harness")
    so = os.path.join(hdir, "libharness.so")
    src = os.path.join(hdir, "harness.cpp")
    deps = [src, os.path.join(ROOT, "pinnacle_b200", "csrc", "jet_math.cuh"), os.path.join(ROOT, "pinnacle_b200", "csrc", "pinn_configs.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", so, src], check=True)
    return ctypes.CDLL(so)
