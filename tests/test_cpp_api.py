"""include/zkb.hpp, the header-only C++ mirror of the reference's consumer API (Source / GpuBackend / Evaluator /
Validator / Stats), compiled with g++ against libzkb.so and run on a golden workspace."""
import os
import subprocess

import pytest

from tests.util import ROOT, has_gpu

SRC = os.path.join(ROOT, "tests", "native", "cpp_api.cpp")
BIN = os.path.join(ROOT, "tests", "native", "cpp_api")
PKG = os.path.join(ROOT, "zkinterface-ir_b200")
WS = os.path.join(ROOT, "tests", "golden", "example")


def build():
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(
            os.path.getmtime(SRC), os.path.getmtime(os.path.join(ROOT, "include", "zkb.hpp")), os.path.getmtime(os.path.join(ROOT, "include", "zkb.h"))):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", BIN,
                               "-L", PKG, "-lzkb", f"-Wl,-rpath,{PKG}"])


def test_cpp_api_host_side():
    build()
    r = subprocess.run([BIN, WS], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "cpp_api ok"


@pytest.mark.gpu
def test_cpp_api_on_device():
    build()
    r = subprocess.run([BIN, WS, "gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "cpp_api ok (gpu)"
