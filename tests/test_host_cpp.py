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


MULTI = os.path.join(ROOT, "tests", "host_cpp", "test_multi")


def build_multi(libdir, libname):
    src = os.path.join(ROOT, "tests", "host_cpp", "test_multi.cpp")
    exe = MULTI + "_" + libname
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-o", exe, src, "-L" + libdir, "-l" + libname, "-Wl,-rpath," + libdir])
    return exe


def test_multi_device_host_calls_from_cpp_under_emulation(emu):
    """The C++ caller's view of the multi-device path (no torch, no IPC, one process), linked against the TEST-ONLY
    emulation pretending to have 4 devices: scatter, slab programs, fused exchange, gather and batch sharding."""
    exe = build_multi(os.path.join(ROOT, "tests", "emu"), "nrb_emu")
    out = subprocess.run([exe, "16", "32", "8"], capture_output=True, text=True, env=dict(os.environ, NRB_EMU_DEVICES="4"))
    assert out.returncode == 0 and "ok (4 device(s))" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_multi_device_host_calls_from_cpp(gpu):
    """On the GPU box: every visible device (one on a 1-GPU box, where both settings take the same path)."""
    exe = build_multi(os.path.join(ROOT, "numrs_b200"), "numrs_b200")
    out = subprocess.run([exe, "64", "128", "64"], capture_output=True, text=True)
    assert out.returncode == 0 and "multi-device host calls: ok" in out.stdout, out.stdout + out.stderr
