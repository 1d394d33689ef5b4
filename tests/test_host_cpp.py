"""C++ host mirror of the reference interface (numrs_b200/host/num_rs.hpp) over the C ABI."""
import os
import subprocess

import pytest

import numrs_b200 as nb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "host_cpp", "test_host")


def build():
    src = os.path.join(ROOT, "tests", "host_cpp", "test_host.cpp")
    libdir = os.path.join(ROOT, "numrs_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-o", EXE, src, "-L" + libdir, "-lnumrs_b200",
                           "-Wl,-rpath," + libdir])


def test_host_mirror_argument_errors_without_device():
    assert os.path.exists(nb.LIB_PATH)
    build()
    out = subprocess.run([EXE], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_host_mirror_reference_unit_tests(gpu):
    build()
    out = subprocess.run([EXE, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
