"""CPU tier, property-based: random shapes / batches / option settings through the C ABI on the TEST-ONLY emulator
(tests/emu) against the oracle.  The hand-picked cases of test_emu_parity.py follow the reference's unit tests; this file
looks for planner corner cases nobody thought of (degenerate dimensions, mixed batch lengths, every factorisation the
tunables allow).  Tolerance: rel-L2 <= 1e-12 * log2 N (BASELINE.json)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import cases
import numrs_b200 as nb
import oracle as O

SET = settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow],
               derandomize=True)


@st.composite
def pow2_shapes(draw, max_ndim=4, max_log2_total=13):
    ndim = draw(st.integers(1, max_ndim))
    logs, left = [], max_log2_total
    for _ in range(ndim):
        lg = draw(st.integers(1, max(1, min(left - (ndim - len(logs) - 1), 11))))
        logs.append(lg)
        left -= lg
    return tuple(1 << lg for lg in logs)


@SET
@given(shape=pow2_shapes(), row_max=st.integers(2, 13), col_max=st.integers(1, 10), isign=st.sampled_from([1, -1]),
       simple=st.integers(0, 1))
def test_fourn_any_shape_any_factorisation(emu, shape, row_max, col_max, isign, simple):
    emu.set_option("row_max_log2", row_max)
    emu.set_option("col_max_log2", col_max)
    emu.set_option("simple_addr", simple)
    n = int(np.prod(shape))
    x = cases.gen(7 + n, 2 * n)
    ref = O.fourn(x.copy(), list(shape), isign)
    got = x.copy()
    nb.fourn(got, list(shape), len(shape), isign, emu)
    assert cases.rel(got, ref) <= cases.tol(n), (shape, row_max, col_max, isign)


@SET
@given(lengths=st.lists(st.sampled_from([1, 2, 4, 8, 32, 128, 512, 2048]), min_size=1, max_size=9), isign=st.sampled_from([1, -1]),
       big=st.integers(0, 1))
def test_fft_batch_mixed_lengths(emu, lengths, isign, big):
    emu.set_option("big_row_mask", (1 << 11) if big else 0)
    arrs = [cases.gen(100 + i, 2 * n) for i, n in enumerate(lengths)]
    refs = [O.four1(a.copy(), a.size // 2, isign) if a.size > 2 else a.copy() for a in arrs]
    nb.FFTProcessor(emu).fft_batch(arrs, isign)
    for a, r in zip(arrs, refs):
        assert cases.rel(a, r) <= cases.tol(a.size)


@SET
@given(lg=st.integers(1, 14), isign=st.sampled_from([1, -1, 0, 5]), row_max=st.integers(3, 13), col_max=st.integers(2, 10))
def test_realft_any_length_any_isign(emu, lg, isign, row_max, col_max):
    # Real_FT.rs:10,15: isign == 1 is the forward transform, ANY other value the inverse
    emu.set_option("row_max_log2", row_max)
    emu.set_option("col_max_log2", col_max)
    n = 1 << lg
    x = cases.gen(300 + lg, n)
    ref = O.realft(x.copy(), n, 1 if isign == 1 else -1)
    got = x.copy()
    nb.realft(got, n, isign, emu)
    assert cases.rel(got, ref) <= cases.tol(n)


@SET
@given(l1=st.integers(0, 5), l2=st.integers(0, 5), l3=st.integers(1, 6), isign=st.sampled_from([1, -1]), side=st.integers(0, 1),
       col_max=st.integers(1, 10))
def test_rlft3_any_shape(emu, l1, l2, l3, isign, side, col_max):
    emu.set_option("speq_side", side)
    emu.set_option("col_max_log2", col_max)
    shp = (1 << l1, 1 << l2, 1 << l3)
    n = int(np.prod(shp))
    x = cases.gen(400 + n, n).reshape(shp)
    s = cases.gen(401 + n, shp[0] * 2 * shp[1]).reshape(shp[0], 2 * shp[1])
    rd, rs = O.rlft3(x.copy(), s.copy(), isign)
    d, sp = x.copy(), s.copy()
    nb.rlft3(d, sp, shp[0], shp[1], shp[2], isign, emu)
    assert cases.rel(d, rd) <= cases.tol(n)
    if isign == 1:
        assert cases.rel(sp, rs) <= cases.tol(n)


@SET
@given(lg=st.integers(1, 13), m_frac=st.floats(0.0, 1.0), isign=st.sampled_from([1, -1]), pad=st.sampled_from([0, 1]),
       count=st.integers(1, 3), transposed=st.integers(0, 1), row_max=st.integers(4, 13), fused=st.integers(0, 1))
def test_convlv_any_response_length(emu, lg, m_frac, isign, pad, count, transposed, row_max, fused):
    emu.set_option("conv_transposed", transposed)
    emu.set_option("conv_fused_mid", fused)
    emu.set_option("row_max_log2", row_max)
    n = 1 << lg
    m = max(1, min(n, int(round(m_frac * n))))
    sigs = [cases.gen(500 + b, n) for b in range(count)]
    # a well-conditioned response for the deconvolution (Convolve.rs:118-122 zeroes bins with |R|^2 < 1e-12)
    r = 0.5 ** np.arange(m) if isign == -1 else cases.gen(600, m) / 8
    refs = [O.convlv(a, r, isign, pad)[1] for a in sigs]
    got = nb.convlv_batch(sigs, r, isign, pad, emu)
    for g, ref in zip(got, refs):
        assert cases.rel(g, ref) <= 50 * cases.tol(n), (n, m, isign, pad)


@SET
@given(n=st.one_of(st.integers(1, 40), st.sampled_from([64, 256, 1024, 4096])), count=st.integers(1, 3),
       row_max=st.sampled_from([4, 6, 13]), fused=st.integers(0, 1))
def test_correl_direct_and_fft_branches(emu, n, count, row_max, fused):
    emu.set_option("row_max_log2", row_max)
    emu.set_option("conv_fused_mid", fused)
    if n > 32 and n & (n - 1):
        with pytest.raises(nb.CorrelError):
            nb.correl(np.ones(n), np.ones(n), emu)          # n > 32 must be a power of two (documented restriction)
        return
    pairs = [(cases.gen(700 + b, n), cases.gen(800 + b, n)) for b in range(count)]
    refs = [O.correl(a, b)[1] for a, b in pairs]
    got = nb.correl_batch(pairs, emu)
    for g, ref in zip(got, refs):
        assert cases.rel(g, ref) <= cases.tol(max(2, n))
