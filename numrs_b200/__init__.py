"""numrs_b200 -- B200-native (sm_100a) implementation of the numrs FFT hot path.

Host-side mirror of the reference's public interface for this path (same names, argument
meaning and error behaviour as /root/reference/src/{FFT_1,Fourn,Real_FT,Real_FT3,Convolve,
Correlation}.rs), calling libnumrs_b200.so through the C ABI of include/numrs_b200.h.
The transforms run on the GPU only: importing works anywhere, but every compute call fails
loudly (NrbError / ImportError) when the CUDA library or a CUDA device is missing.
"""
import os

import numpy as np

from . import _lib
from ._lib import (NRB_PAD_LITERAL, NRB_PAD_NR, NrbError, KIND_FOUR1, KIND_FOURN, KIND_REALFT, KIND_RLFT3,
                   KIND_CONVLV, KIND_CORREL, KIND_CORREL_NORM, KIND_CORREL_NORM_FAST, KIND_AUTOCORREL_FAST,
                   KIND_TWOFFT, KIND_POWER, KIND_COSFT1, KIND_COSFT2, KIND_SINFT)

_HERE = os.path.dirname(os.path.abspath(__file__))
# NUMRS_B200_LIB may point at another build of the same CUDA library (tuning experiments)
LIB_PATH = os.environ.get("NUMRS_B200_LIB") or os.path.join(_HERE, "libnumrs_b200.so")
_LIB = None


def lib():
    """The loaded CUDA library (numrs_b200/libnumrs_b200.so).  Raises ImportError if not built."""
    global _LIB
    if _LIB is None:
        _LIB = _lib.Library(LIB_PATH)
    return _LIB


# ---------------------------------------------------------------- error types (reference enums)
class ConvlvError(Exception):
    """Convolve.rs:226-238"""
    EmptyInput = "EmptyInput"
    ResponseTooLong = "ResponseTooLong"
    InvalidIsign = "InvalidIsign"
    DivisionByZero = "DivisionByZero"     # never produced, as in the reference (ledger L2)
    FftError = "FftError"

    def __init__(self, kind, detail=""):
        super().__init__(f"{kind}: {detail}" if detail else kind)
        self.kind = kind


class CorrelError(Exception):
    """Correlation.rs:389-399"""
    EmptyInput = "EmptyInput"
    LengthMismatch = "LengthMismatch"
    FftError = "FftError"
    ZeroStdDev = "ZeroStdDev"

    def __init__(self, kind, detail=""):
        super().__init__(f"{kind}: {detail}" if detail else kind)
        self.kind = kind


def _convlv_raise(L, rc):
    kind = {_lib.NRB_ERR_EMPTY_INPUT: ConvlvError.EmptyInput,
            _lib.NRB_ERR_RESPONSE_TOO_LONG: ConvlvError.ResponseTooLong,
            _lib.NRB_ERR_INVALID_ISIGN: ConvlvError.InvalidIsign}.get(rc, ConvlvError.FftError)
    raise ConvlvError(kind, L.last_error())


def _correl_raise(L, rc):
    kind = {_lib.NRB_ERR_EMPTY_INPUT: CorrelError.EmptyInput,
            _lib.NRB_ERR_LENGTH_MISMATCH: CorrelError.LengthMismatch,
            _lib.NRB_ERR_ZERO_STDDEV: CorrelError.ZeroStdDev}.get(rc, CorrelError.FftError)
    raise CorrelError(kind, L.last_error())


def _panic(L, rc):
    """four1 / realft / rlft3 return () in the reference and signal misuse by panicking."""
    if rc != 0:
        raise NrbError(rc, L.last_error())


# ---------------------------------------------------------------- FFT_1.rs
def four1(data, nn, isign, _L=None):
    """FFT_1.rs:5 `four1(data: &mut [f64], nn, isign)`: in place, unnormalised."""
    # the reference indexes data[..2 * nn] and panics on a shorter slice; the C ABI copies exactly 2 * nn doubles
    assert data.size >= 2 * nn, "data length must be at least 2 * nn"
    L = _L or lib()
    _panic(L, L.four1(data, nn, isign))


four1_optimized = four1   # FFT_1.rs:110: same transform (numerically equivalent variant)


class FFTProcessor:
    """FFT_1.rs:143-190.  Builder flags are accepted for API compatibility; there is one GPU path."""

    def __init__(self, _L=None):
        self.max_threads = os.cpu_count() or 1
        self.use_optimized = True
        self._L = _L

    def with_threads(self, threads):
        self.max_threads = threads
        return self

    def with_optimized(self, use_optimized):
        self.use_optimized = use_optimized
        return self

    def fft(self, data, isign):
        four1(data, data.size // 2, isign, self._L)

    def fft_batch(self, batches, isign):
        L = self._L or lib()
        _panic(L, L.four1_batch(list(batches), isign))


def real_to_complex(real_data):          # FFT_1.rs:193-200
    out = np.zeros(2 * len(real_data), dtype=np.float64)
    out[0::2] = real_data
    return out


def complex_to_real(complex_data):       # FFT_1.rs:202-204
    return np.asarray(complex_data, dtype=np.float64)[0::2].copy()


def power_spectrum(complex_data, _L=None):        # FFT_1.rs:218-228 (a trailing odd element is dropped)
    L = _L or lib()
    rc, out = L.power_spectrum(complex_data, False)
    _panic(L, rc)
    return out


def magnitude_spectrum(complex_data, _L=None):    # FFT_1.rs:206-216
    L = _L or lib()
    rc, out = L.power_spectrum(complex_data, True)
    _panic(L, rc)
    return out


# ---------------------------------------------------------------- FFT_2.rs
def twofft(data1, data2, fft1, fft2, _L=None):
    """FFT_2.rs:3 `twofft(data1, data2, fft1, fft2)`: asserts as FFT_2.rs:5-7; NR semantics (0-based mirror n - k)."""
    n = data1.size
    assert data2.size == n, "data2 length must equal data1 length"
    assert fft1.size == 2 * n + 2, "fft1 must have length 2*n + 2"
    assert fft2.size == 2 * n + 2, "fft2 must have length 2*n + 2"
    L = _L or lib()
    _panic(L, L.twofft(data1, data2, fft1, fft2))


def twofft_optimized(data1, data2, fft1, fft2, _L=None):
    """FFT_2.rs:135: same contract as `twofft`."""
    twofft(data1, data2, fft1, fft2, _L)


class TwoFFTProcessor:
    """FFT_2.rs:222-265.  `use_optimized` / `parallel_threshold` select between CPU code paths in the reference;
    they are kept and ignored (one device path)."""

    def __init__(self, _L=None):
        self.use_optimized = True          # FFT_2.rs:230
        self.parallel_threshold = 1024     # FFT_2.rs:231
        self._L = _L

    def with_optimized(self, use_optimized):
        self.use_optimized = bool(use_optimized)
        return self

    def with_threshold(self, threshold):
        self.parallel_threshold = int(threshold)
        return self

    def process(self, data1, data2, fft1, fft2):
        twofft(data1, data2, fft1, fft2, self._L)

    def process_batch(self, batches):
        """FFT_2.rs:258 `process_batch(&[(data1, data2, fft1, fft2)])`: tuples of equal length share one batched
        device plan (the reference loops over them)."""
        L = self._L or lib()
        groups = {}
        for d1, d2, f1, f2 in batches:
            n = d1.size
            assert d2.size == n, "data2 length must equal data1 length"             # FFT_2.rs:5
            assert f1.size == 2 * n + 2, "fft1 must have length 2*n + 2"            # FFT_2.rs:6
            assert f2.size == 2 * n + 2, "fft2 must have length 2*n + 2"            # FFT_2.rs:7
            groups.setdefault(n, []).append((d1, d2, f1, f2))
        for n, items in groups.items():
            _panic(L, L.twofft_batch([i[0] for i in items], [i[1] for i in items], [i[2] for i in items], [i[3] for i in items]))


def extract_real_imag(fft):                 # FFT_2.rs:361-372
    fft = np.asarray(fft, dtype=np.float64)
    n = fft.size // 2
    return fft[0:2 * n:2].copy(), fft[1:2 * n:2].copy()


def combine_real_imag(real, imag):          # FFT_2.rs:375-385
    real = np.asarray(real, dtype=np.float64)
    imag = np.asarray(imag, dtype=np.float64)
    assert real.size == imag.size, "Real and imaginary parts must have same length"
    out = np.empty(2 * real.size, dtype=np.float64)
    out[0::2] = real
    out[1::2] = imag
    return out


# ---------------------------------------------------------------- Cos_FT.rs / Cos_FT2.rs / sinft
def cosft1(y, n, _L=None):
    """Cos_FT.rs:7 `cosft1(y: &mut [f64], n)`: 1-based array, y[0] unused, data y[1..=n+1]; NR semantics."""
    assert y.size >= n + 2, "y must hold n + 2 elements"        # Cos_FT.rs:17 reads y[n + 1]
    L = _L or lib()
    _panic(L, L.cosft1(y, n))


def cosft1_optimized(y, n, _L=None):
    """Cos_FT.rs:78: same contract as `cosft1`."""
    cosft1(y, n, _L)


def cosft2(y, n, isign, _L=None):
    """Cos_FT2.rs:7 `cosft2(y, n, isign)`: 1-based array, data y[1..=n]; panics on isign not in {1, -1}."""
    assert y.size >= n + 1, "y must hold n + 1 elements"
    L = _L or lib()
    _panic(L, L.cosft2(y, n, isign))


def cosft2_simd(y, n, isign, _L=None):
    """Cos_FT2.rs:202: same contract as `cosft2`."""
    cosft2(y, n, isign, _L)


def sinft(y, n, _L=None):
    """NR sinft (README.md:72): 1-based array, data y[1..=n], y[1] is taken as 0."""
    assert y.size >= n + 1, "y must hold n + 1 elements"
    L = _L or lib()
    _panic(L, L.sinft(y, n))


# ---------------------------------------------------------------- Fourn.rs / Real_FT3.rs:35
def fourn(data, nn, ndim, isign, _L=None):
    """In-memory N-dimensional complex FFT with the call shape of Real_FT3.rs:35 and the
    validation of Fourn.rs:367-378 (raises ValueError like io::ErrorKind::InvalidInput)."""
    L = _L or lib()
    nn = list(nn)
    if ndim == 0 or ndim > len(nn):
        raise ValueError("Invalid dimensions")
    total = 1
    for d in nn[:ndim]:
        total *= int(d)
    # an in-memory fourn indexes data[..2 * prod(nn)] (panic on a shorter slice); the C ABI copies exactly that many
    assert total <= 0 or data.size >= 2 * total, "data length must be at least 2 * prod(nn)"
    rc = L.fourn(data, nn, ndim, isign)
    if rc in (_lib.NRB_ERR_INVALID_DIMS, _lib.NRB_ERR_INVALID_ISIGN):
        raise ValueError(L.last_error())
    _panic(L, rc)


# ---------------------------------------------------------------- Real_FT.rs
def realft(data, n, isign, _L=None):
    """Real_FT.rs:4 `realft(data, n, isign)`; asserts as Real_FT.rs:5-6."""
    assert n % 2 == 0, "n must be even"
    assert data.size >= n, "data length must be at least n"
    L = _L or lib()
    _panic(L, L.realft(data, n, isign))


realft_optimized = realft


class RealFTProcessor:
    """Real_FT.rs:332-370"""

    def __init__(self, _L=None):
        self.use_optimized = True
        self.parallel_threshold = 1024
        self._L = _L

    def with_optimized(self, use_optimized):
        self.use_optimized = use_optimized
        return self

    def with_threshold(self, threshold):
        self.parallel_threshold = threshold
        return self

    def process(self, data, n, isign):
        realft(data, n, isign, self._L)

    def process_batch(self, batches):
        """batches: list of (data, n, isign) as in Real_FT.rs:365-369; equal (n, isign) runs are batched."""
        L = self._L or lib()
        groups = {}
        for data, n, isign in batches:
            assert n % 2 == 0 and data.size >= n
            groups.setdefault((n, 1 if isign == 1 else -1), []).append(data)
        for (n, isign), arrs in groups.items():
            _panic(L, L.realft_batch(arrs, n, isign))


# ---------------------------------------------------------------- Real_FT3.rs
def rlft3(data, speq, nn1, nn2, nn3, isign, _L=None):
    """Real_FT3.rs:8 `rlft3(data: Array3, speq: Array2, nn1, nn2, nn3, isign)`; asserts :17-19."""
    assert isign == 1 or isign == -1, "isign must be 1 or -1"
    assert data.shape == (nn1, nn2, nn3), "data dimensions mismatch"
    assert speq.shape == (nn1, 2 * nn2), "speq dimensions mismatch"
    L = _L or lib()
    _panic(L, L.rlft3(data, speq, nn1, nn2, nn3, isign))


def rlft3_optimized(data, speq, nn1, nn2, nn3, isign, _L=None):
    """Real_FT3.rs:145: flat-slice variant, same result (ledger D5)."""
    assert isign in (1, -1), "isign must be 1 or -1"
    assert data.size == nn1 * nn2 * nn3, "data dimensions mismatch"
    assert speq.size == 2 * nn1 * nn2, "speq dimensions mismatch"
    # in place: a non-contiguous view would be transformed in a temporary copy and the result lost
    assert data.flags["C_CONTIGUOUS"] and speq.flags["C_CONTIGUOUS"], "data and speq must be C-contiguous (as_slice_mut, Real_FT3.rs:33)"
    L = _L or lib()
    _panic(L, L.rlft3(data.reshape(-1), speq.reshape(-1), nn1, nn2, nn3, isign))


# ---------------------------------------------------------------- Convolve.rs
def convlv(data, respns, isign, pad_mode=NRB_PAD_LITERAL, _L=None):
    """Convolve.rs:8: returns the length-n result, raises ConvlvError like the reference's Err."""
    L = _L or lib()
    rc, ans = L.convlv(data, respns, isign, pad_mode)
    if rc != 0:
        _convlv_raise(L, rc)
    return ans


def convlv_batch(data_batch, respns, isign, pad_mode=NRB_PAD_LITERAL, _L=None):
    """Convolve.rs:241; signals of equal length are transformed as one device batch."""
    L = _L or lib()
    sigs = [np.ascontiguousarray(d, dtype=np.float64) for d in data_batch]
    if len({s.size for s in sigs}) > 1:
        return [convlv(s, respns, isign, pad_mode, L) for s in sigs]
    if not sigs:
        return []
    if sigs[0].size == 0:
        raise ConvlvError(ConvlvError.EmptyInput)
    rc, outs = L.convlv_batch(sigs, respns, isign, pad_mode)
    if rc != 0:
        _convlv_raise(L, rc)
    return outs


class ConvlvProcessor:
    """Convolve.rs:253-339"""

    def __init__(self, _L=None):
        self.use_optimized = True
        self.parallel_threshold = 1024
        self._L = _L

    def with_optimized(self, use_optimized):
        self.use_optimized = use_optimized
        return self

    def with_threshold(self, threshold):
        self.parallel_threshold = threshold
        return self

    def process(self, data, respns, isign):
        return convlv(data, respns, isign, _L=self._L)

    def process_batch(self, data_batch, respns, isign):
        return convlv_batch(data_batch, respns, isign, _L=self._L)


# ---------------------------------------------------------------- Correlation.rs
def correl(data1, data2, _L=None):
    """Correlation.rs:8 (n <= 32: linear direct lags, n > 32: circular via realft)."""
    L = _L or lib()
    rc, ans = L.correl(data1, data2)
    if rc != 0:
        _correl_raise(L, rc)
    return ans


def correl_batch(data_pairs, _L=None):
    """Correlation.rs:273"""
    L = _L or lib()
    pairs = [(np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64))
             for a, b in data_pairs]
    if not pairs:
        return []
    sizes = {(a.size, b.size) for a, b in pairs}
    if len(sizes) > 1 or any(a.size != b.size or a.size == 0 for a, b in pairs):
        return [correl(a, b, L) for a, b in pairs]
    rc, outs = L.correl_batch([a for a, _ in pairs], [b for _, b in pairs])
    if rc != 0:
        _correl_raise(L, rc)
    return outs


def autocorrel(data, _L=None):
    """Correlation.rs:281"""
    return correl(data, data, _L)


def correl_normalized(data1, data2, _L=None):
    """Correlation.rs:189: two-pass population statistics, ZeroStdDev, correl of the normalised signals."""
    L = _L or lib()
    rc, ans = L.correl_normalized(data1, data2, False)
    if rc != 0:
        _correl_raise(L, rc)
    return ans


def correl_normalized_fast(data1, data2, _L=None):
    """Correlation.rs:226: single-pass statistics; n <= 32 direct lags scaled by 1/(std1 std2 n)."""
    L = _L or lib()
    rc, ans = L.correl_normalized(data1, data2, True)
    if rc != 0:
        _correl_raise(L, rc)
    return ans


def autocorrel_fast(data, _L=None):
    """Correlation.rs:286: n <= 32 direct lags; else forward realft, |F|^2, inverse realft (ledger D10)."""
    L = _L or lib()
    rc, ans = L.autocorrel_fast(data)
    if rc != 0:
        _correl_raise(L, rc)
    return ans
