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


def _c_prototypes():
    """name -> number of parameters, from include/numrs_b200.h"""
    hdr = open(os.path.join(ROOT, "include", "numrs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for name, args in re.findall(r"\b(nrb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr):
        args = args.strip()
        protos[name] = 0 if args in ("", "void") else len(args.split(","))
    return protos


def test_rust_shim_declarations_match_the_header():
    """rust/src/ffi.rs cannot be compiled here (no cargo): check it textually against the C header -- every extern fn it
    declares exists in the header with the same number of parameters, every constant has the header's value, and every
    host-slice entry point a reference function binds to is declared."""
    src = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    ext = src[src.index('extern "C" {'):]
    ext = ext[:ext.index("\n}\n")]
    ext = re.sub(r"//[^\n]*", "", ext)
    rust = {}
    for name, args in re.findall(r"pub fn (nrb_[a-z0-9_]+)\s*\(([^)]*)\)", ext, flags=re.S):
        args = args.strip()
        rust[name] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    protos = _c_prototypes()
    assert len(rust) >= 35
    for name, nargs in rust.items():
        assert name in protos, f"{name} is declared in ffi.rs but not in the header"
        assert nargs == protos[name], (name, nargs, protos[name])
    for name in ("nrb_four1", "nrb_four1_batch", "nrb_fourn", "nrb_realft", "nrb_realft_batch", "nrb_rlft3", "nrb_convlv",
                 "nrb_convlv_batch", "nrb_correl", "nrb_correl_batch", "nrb_correl_normalized", "nrb_autocorrel_fast",
                 "nrb_twofft", "nrb_twofft_batch", "nrb_cosft1", "nrb_cosft2", "nrb_sinft", "nrb_power_spectrum",
                 "nrb_set_option", "nrb_last_error", "nrb_plan_create", "nrb_plan_exec", "nrb_plan_destroy"):
        assert name in rust, name
    hdr = open(os.path.join(ROOT, "include", "numrs_b200.h")).read()
    cdefs = {k: int(v) for k, v in re.findall(r"#define\s+(NRB_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", hdr)}
    for k, v in re.findall(r"pub const (NRB_[A-Z0-9_]+): c_int = (-?\d+);", src):
        assert cdefs.get(k) == int(v), (k, v, cdefs.get(k))
    # every nrb_* call in the shim's modules is declared in ffi.rs
    lib_rs = open(os.path.join(ROOT, "rust", "src", "lib.rs")).read()
    for name in set(re.findall(r"\b(nrb_[a-z0-9_]+)\s*\(", lib_rs)):
        assert name in rust, f"rust/src/lib.rs calls {name}, which ffi.rs does not declare"


def test_every_option_is_documented_in_the_header():
    """nrb_set_option's names (plan.cpp set_tunable) all appear in the option list of include/numrs_b200.h"""
    src = open(os.path.join(ROOT, "numrs_b200", "csrc", "plan.cpp")).read()
    names = set(re.findall(r'n == "([a-z0-9_]+)"', src))
    assert len(names) >= 25
    hdr = open(os.path.join(ROOT, "include", "numrs_b200.h")).read()
    missing = sorted(n for n in names if f'"{n}"' not in hdr)
    assert not missing, missing


def test_headline_kernels_keep_their_register_budget():
    """The occupancy the measured numbers rest on is a property of the build: the z pass of rlft3 512^3
    (ROW-REAL, 256 complex points) and the register-fed strided pass run with <= 64 registers (2 CTAs of 512 threads or 4+ of
    256 per SM), the TMA-fed 512-point strided pass with <= 85 (three 256-thread CTAs per SM), the fused conv middle with
    <= 64 (one 1024-thread CTA), and none of them has a local-memory frame worth a spill."""
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--dump-resource-usage", nb.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    items = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out.stdout)
    assert len(items) > 100
    dem = subprocess.run(["c++filt"], input="\n".join(i[0] for i in items), capture_output=True, text=True).stdout.split("\n")
    usage = {d: (int(r), int(s), int(l)) for d, (_, r, s, _, l) in zip(dem, items)}

    def find(prefix):
        hits = [v for k, v in usage.items() if k.replace("void nrb::", "").startswith(prefix)]
        assert hits, prefix
        return hits
    budget = {
        "fft_pass_kernel<8, 0, 1, 1>": 64, "fft_pass_kernel<8, 0, -1, 1>": 64,          # ROW-REAL 256 (z pass)
        "fft_pass_kernel<9, 1, 1, 0>": 128, "fft_pass_kernel<9, 1, -1, 0>": 128,        # strided 512, register-fed (16 points / thread)
        "fft_col_tma_kernel<9, 1, false>": 85, "fft_col_tma_kernel<9, -1, false>": 85,  # strided 512, TMA-fed, 3 CTAs / SM
        "fft_col_tma_kernel<10, 1, false>": 85, "fft_col_tma_kernel<10, -1, false>": 85,
        "conv_mid_kernel<12>": 64,
        "trig_kernel<11>": 64, "twofft_kernel<12>": 64,
    }
    for prefix, regs in budget.items():
        for r, stack, local in find(prefix):
            assert r <= regs, (prefix, r, regs)
            assert stack <= 16 and local == 0, (prefix, stack, local)
