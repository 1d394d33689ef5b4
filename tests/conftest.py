import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def emu():
    """TEST-ONLY host emulation of the kernels (tests/emu); validates planner + index math on CPU."""
    from numrs_b200 import _lib
    d = os.path.join(ROOT, "tests", "emu")
    # one build at a time: under pytest-xdist every worker runs this fixture, and a worker must not load a library that
    # another one is still linking
    import fcntl
    with open(os.path.join(d, ".build.lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        subprocess.check_call(["make", "-C", d, "-s"])
    return _lib.Library(os.path.join(d, "libnrb_emu.so"))


@pytest.fixture(scope="session")
def gpu():
    """The product CUDA library on a real device.  No device => hard failure (no CPU fallback)."""
    import numrs_b200
    L = numrs_b200.lib()
    if L.device_count() <= 0:
        pytest.fail("no CUDA device visible: numrs_b200 has no CPU fallback")
    return L


@pytest.fixture(autouse=True)
def _reset_options(request):
    yield
    for name in ("emu", "gpu"):
        if name in request.fixturenames:
            L = request.getfixturevalue(name)
            L.set_option("col_max_log2", 10)
            L.set_option("row_max_log2", 13)
            L.set_option("l2_group_bytes", 1 << 40)
            L.set_option("batch_group_bytes", 512 << 20)
            L.set_option("fuse_zy", 0)
            L.set_option("fuse_lag", 16)
            L.set_option("prefetch_dist", -1)
            L.set_option("conv_transposed", 1)
            L.set_option("xchg_grid_cap", 0)
            L.set_option("simple_addr", 1)
            L.set_option("speq_side", 1)
            L.set_option("conv_fused_mid", 1)
            L.set_option("big_row_mask", 0)
            L.set_option("big_col_mask", 0)
            L.set_option("conv_rest_log2", 12)
            L.set_option("mid_prefetch", 0)
            L.set_option("pull_eighths", 4)
            L.set_option("z_chunks", 1)
            L.set_option("pipeline_batches", 1)
            L.set_option("pipeline_min_kb", 16384)
            L.set_option("dma_streams", 1)
            L.set_option("tma_col_mask", (1 << 9) | (1 << 10))
            L.set_option("tma_persist", 0)
            L.set_option("tma_xpose", 1)
            L.set_option("tma_in_mask", 0)
            L.set_option("tma_in_ctas", 2)
            L.set_option("trig_fused", 1)
