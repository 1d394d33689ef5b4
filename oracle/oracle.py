"""ctypes loader for oracle/liboracle.so (nr_oracle.c).  Test infrastructure only.

Every wrapper works on numpy float64 arrays using the reference's layouts: interleaved complex
(`data[2k]=Re, data[2k+1]=Im`), in-place transforms, unnormalised inverses.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_sz = ctypes.c_size_t


def build(force=False):
    """Compile nr_oracle.c -> liboracle.so (gcc; seconds)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "nr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = ctypes.CDLL(so)
        L.orc_fill_uniform.argtypes = [ctypes.c_uint64, ctypes.c_uint64, _sz, _dp]
        L.orc_fill_uniform.restype = None
        for name in ("orc_four1", "orc_four1_mt", "orc_four1_optimized"):
            f = getattr(L, name)
            f.argtypes = [_dp, _sz, ctypes.c_int]
            f.restype = None
        L.orc_fft_batch.argtypes = [ctypes.POINTER(_dp), ctypes.POINTER(_sz), _sz, ctypes.c_int, ctypes.c_int]
        L.orc_fft_batch.restype = None
        for name in ("orc_fourn", "orc_fourn_mt"):
            f = getattr(L, name)
            f.argtypes = [_dp, ctypes.POINTER(_sz), ctypes.c_int, ctypes.c_int]
            f.restype = None
        L.orc_fourn_validate.argtypes = [ctypes.POINTER(_sz), _sz, _sz, ctypes.c_int]
        L.orc_fourn_validate.restype = ctypes.c_int
        L.orc_realft.argtypes = [_dp, _sz, ctypes.c_int]
        L.orc_realft.restype = None
        for name in ("orc_rlft3", "orc_rlft3_mt"):
            f = getattr(L, name)
            f.argtypes = [_dp, _dp, _sz, _sz, _sz, ctypes.c_int]
            f.restype = None
        L.orc_pad_response.argtypes = [_dp, _sz, _sz, ctypes.c_int, _dp]
        L.orc_pad_response.restype = None
        L.orc_convlv.argtypes = [_dp, _sz, _dp, _sz, ctypes.c_int, ctypes.c_int, _dp]
        L.orc_convlv.restype = ctypes.c_int
        L.orc_convlv_batch.argtypes = [ctypes.POINTER(_dp), _sz, _sz, _dp, _sz, ctypes.c_int, ctypes.c_int,
                                       ctypes.POINTER(_dp), ctypes.c_int]
        L.orc_convlv_batch.restype = ctypes.c_int
        L.orc_correl.argtypes = [_dp, _sz, _dp, _sz, _dp]
        L.orc_correl.restype = ctypes.c_int
        L.orc_correl_batch.argtypes = [ctypes.POINTER(_dp), ctypes.POINTER(_dp), _sz, _sz, ctypes.POINTER(_dp),
                                       ctypes.c_int]
        L.orc_correl_batch.restype = ctypes.c_int
        L.orc_twofft.argtypes = [_dp, _dp, _sz, _dp, _dp]
        L.orc_twofft.restype = None
        L.orc_power_spectrum.argtypes = [_dp, _sz, ctypes.c_int, _dp]
        L.orc_power_spectrum.restype = None
        L.orc_correl_normalized.argtypes = [_dp, _sz, _dp, _sz, ctypes.c_int, _dp]
        L.orc_correl_normalized.restype = ctypes.c_int
        L.orc_autocorrel_fast.argtypes = [_dp, _sz, _dp]
        L.orc_autocorrel_fast.restype = ctypes.c_int
        L.orc_cosft1.argtypes = [_dp, _sz]
        L.orc_cosft1.restype = None
        L.orc_cosft2.argtypes = [_dp, _sz, ctypes.c_int]
        L.orc_cosft2.restype = ctypes.c_int
        L.orc_sinft.argtypes = [_dp, _sz]
        L.orc_sinft.restype = None
        L.orc_num_threads.argtypes = []
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_set_num_threads.argtypes = [ctypes.c_int]
        L.orc_set_num_threads.restype = None
        _LIB = L
    return _LIB


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def fill_uniform(seed, offset, count):
    """SURVEY.md 8d generator: iid uniform [-1,1) from splitmix64(seed*phi + offset + i)."""
    out = np.empty(count, dtype=np.float64)
    lib().orc_fill_uniform(seed, offset, count, _p(out))
    return out


def fill_uniform_py(seed, offset, count):
    """Pure-numpy restatement of the generator (cross-checks the C / CUDA ones)."""
    M = (1 << 64) - 1
    with np.errstate(over="ignore"):
        x = (np.uint64((seed * 0x9E3779B97F4A7C15 + offset) & M) + np.arange(count, dtype=np.uint64))
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 4503599627370496.0) - 1.0


def four1(data, nn, isign, mt=False):
    (lib().orc_four1_mt if mt else lib().orc_four1)(_p(data), nn, isign)
    return data


def four1_optimized(data, nn, isign):
    lib().orc_four1_optimized(_p(data), nn, isign)
    return data


def fft_batch(arrays, isign, mt=False):
    n = len(arrays)
    ptrs = (_dp * n)(*[_p(a) for a in arrays])
    nn = (_sz * n)(*[a.size // 2 for a in arrays])
    lib().orc_fft_batch(ptrs, nn, n, isign, int(mt))
    return arrays


def fourn(data, nn, isign, mt=False):
    nn_c = (_sz * len(nn))(*nn)
    (lib().orc_fourn_mt if mt else lib().orc_fourn)(_p(data), nn_c, len(nn), isign)
    return data


def fourn_validate(nn, ndim, isign):
    nn_c = (_sz * max(1, len(nn)))(*nn)
    return lib().orc_fourn_validate(nn_c, len(nn), ndim, isign)


def realft(data, n, isign):
    assert n % 2 == 0 and data.size >= n          # Real_FT.rs:5-6
    lib().orc_realft(_p(data), n, isign)
    return data


def rlft3(data, speq, isign, mt=False):
    nn1, nn2, nn3 = data.shape
    assert isign in (1, -1) and speq.shape == (nn1, 2 * nn2)   # Real_FT3.rs:17-19
    (lib().orc_rlft3_mt if mt else lib().orc_rlft3)(_p(data), _p(speq), nn1, nn2, nn3, isign)
    return data, speq


def pad_response(respns, n, pad_mode=0):
    respns = np.ascontiguousarray(respns, dtype=np.float64)
    out = np.empty(n, dtype=np.float64)
    lib().orc_pad_response(_p(respns), respns.size, n, pad_mode, _p(out))
    return out


def convlv(data, respns, isign, pad_mode=0):
    """Returns (rc, ans).  rc follows include/numrs_b200.h error codes."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    respns = np.ascontiguousarray(respns, dtype=np.float64)
    ans = np.zeros(max(1, data.size), dtype=np.float64)
    rc = lib().orc_convlv(_p(data), data.size, _p(respns), respns.size, isign, pad_mode, _p(ans))
    return rc, ans[:data.size]


def convlv_batch(signals, respns, isign, pad_mode=0, mt=True):
    n = signals[0].size
    cnt = len(signals)
    respns = np.ascontiguousarray(respns, dtype=np.float64)
    outs = [np.zeros(n, dtype=np.float64) for _ in range(cnt)]
    ip = (_dp * cnt)(*[_p(s) for s in signals])
    op = (_dp * cnt)(*[_p(o) for o in outs])
    rc = lib().orc_convlv_batch(ip, cnt, n, _p(respns), respns.size, isign, pad_mode, op, int(mt))
    return rc, outs


def correl(d1, d2):
    d1 = np.ascontiguousarray(d1, dtype=np.float64)
    d2 = np.ascontiguousarray(d2, dtype=np.float64)
    ans = np.zeros(max(1, d1.size), dtype=np.float64)
    rc = lib().orc_correl(_p(d1), d1.size, _p(d2), d2.size, _p(ans))
    return rc, ans[:d1.size]


def correl_batch(a_list, b_list, mt=True):
    n = a_list[0].size
    cnt = len(a_list)
    outs = [np.zeros(n, dtype=np.float64) for _ in range(cnt)]
    ap = (_dp * cnt)(*[_p(a) for a in a_list])
    bp = (_dp * cnt)(*[_p(b) for b in b_list])
    op = (_dp * cnt)(*[_p(o) for o in outs])
    rc = lib().orc_correl_batch(ap, bp, cnt, n, op, int(mt))
    return rc, outs


def twofft(d1, d2):
    """FFT_2.rs:3: returns (fft1, fft2), each 2n+2 doubles (n complex bins + 2 zero pad doubles)."""
    d1 = np.ascontiguousarray(d1, dtype=np.float64)
    d2 = np.ascontiguousarray(d2, dtype=np.float64)
    assert d1.size == d2.size
    f1 = np.zeros(2 * d1.size + 2)
    f2 = np.zeros(2 * d1.size + 2)
    lib().orc_twofft(_p(d1), _p(d2), d1.size, _p(f1), _p(f2))
    return f1, f2


def power_spectrum(c, take_sqrt=False):
    c = np.ascontiguousarray(c, dtype=np.float64)
    out = np.zeros(c.size // 2)
    lib().orc_power_spectrum(_p(c), c.size // 2, int(take_sqrt), _p(out))
    return out


def correl_normalized(d1, d2, fast=False):
    d1 = np.ascontiguousarray(d1, dtype=np.float64)
    d2 = np.ascontiguousarray(d2, dtype=np.float64)
    ans = np.zeros(max(1, d1.size))
    rc = lib().orc_correl_normalized(_p(d1), d1.size, _p(d2), d2.size, int(fast), _p(ans))
    return rc, ans[:d1.size]


def autocorrel_fast(d):
    d = np.ascontiguousarray(d, dtype=np.float64)
    ans = np.zeros(max(1, d.size))
    rc = lib().orc_autocorrel_fast(_p(d), d.size, _p(ans))
    return rc, ans[:d.size]


def cosft1(y, n):
    """Cos_FT.rs:7: in place on the 1-based array y[0..n+2) (y[0] unused, data y[1..=n+1])."""
    assert y.dtype == np.float64 and y.size >= n + 2
    lib().orc_cosft1(_p(y), n)
    return y


def cosft2(y, n, isign):
    """Cos_FT2.rs:7: in place on the 1-based array y[0..n+1) (y[0] unused, data y[1..=n])."""
    assert y.dtype == np.float64 and y.size >= n + 1
    rc = lib().orc_cosft2(_p(y), n, isign)
    return rc, y


def sinft(y, n):
    """NR sinft (README.md:72): in place on the 1-based array y[0..n+1)."""
    assert y.dtype == np.float64 and y.size >= n + 1
    lib().orc_sinft(_p(y), n)
    return y


def num_threads():
    return lib().orc_num_threads()


def use_all_cores():
    """Use every core this process may run on, whatever OMP_NUM_THREADS says (torchrun exports 1); returns the count."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(n)
    return num_threads()
