"""CPU tier: the product library loads, exports every symbol include/numrs_b200.h declares, and
fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

import numrs_b200 as nb
from numrs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "numrs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(nrb_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.ABI_SYMBOLS)


def test_cuda_library_exports_every_declared_symbol():
    assert os.path.exists(nb.LIB_PATH), "numrs_b200/libnumrs_b200.so is not built (python -c 'import __graft_entry__ as g; g.build()')"
    L = ctypes.CDLL(nb.LIB_PATH)
    for sym in declared_symbols():
        assert hasattr(L, sym), sym
    lib = nb.lib()
    assert "sm_100a" in lib.version()


def test_cuda_library_contains_only_sm100a_code():
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", nb.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device():
    lib = nb.lib()
    if lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    x = np.zeros(16)
    assert lib.four1(x, 8, 1) == _lib.NRB_ERR_CUDA
    assert "no CPU fallback" in lib.last_error()
    with pytest.raises(nb.NrbError):
        nb.four1(x, 8, 1)
    with pytest.raises(nb.NrbError):
        lib.plan_create(nb.KIND_FOUR1, [8])
    with pytest.raises(nb.ConvlvError) as ei:
        nb.convlv(np.ones(8), np.ones(2), 1)
    assert ei.value.kind == nb.ConvlvError.FftError
    # argument errors are still reported in the reference's order before touching the device
    with pytest.raises(nb.ConvlvError) as ei:
        nb.convlv([], [1.0], 1)
    assert ei.value.kind == nb.ConvlvError.EmptyInput


def test_product_package_never_imports_oracle_or_emulator():
    for root, _, files in os.walk(os.path.join(ROOT, "numrs_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "liboracle" not in src and "libnrb_emu" not in src, f
