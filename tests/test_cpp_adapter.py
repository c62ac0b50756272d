"""The header-only C++ adapter (include/mfmg_b200/mfmg.hpp): compiles against the C ABI (CPU), and the reference's
device unit tests restated in C++ pass on the GPU (-m gpu)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_adapter.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_adapter")


def _compile():
    cmd = ["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC,
           "-L" + os.path.join(ROOT, "mfmg_b200", "csrc"), "-lmfmg_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "mfmg_b200", "csrc"), "-o", EXE]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout


def test_adapter_compiles_and_links():
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_adapter_reference_tests_on_gpu():
    _compile()
    res = subprocess.run([EXE], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0 and "ALL OK" in res.stdout, res.stdout
