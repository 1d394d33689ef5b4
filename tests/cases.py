"""Parity cases shared by the CPU tier (emulated kernels) and the GPU tier (CUDA library).

Every check compares the library under test, called through the C ABI (host-slice entry
points of include/numrs_b200.h via numrs_b200's mirror of the reference interface), against
the CPU oracle (oracle/nr_oracle.c) on the same seeded input.

Tolerance (BASELINE.json north_star): relative L2 error <= 1e-12 * log2(N).
"""
import math

import numpy as np

import numrs_b200 as nb
import oracle as O


def tol(npoints):
    return 1e-12 * max(1.0, math.log2(max(2, npoints)))


def rel(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    nb_ = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (nb_ if nb_ > 0 else 1.0))


def gen(seed, count, offset=0):
    return O.fill_uniform(seed, offset, count)


def check_four1(L, nn, seed=1001):
    x = gen(seed, 2 * nn)
    for isign in (1, -1):
        ref = O.four1(x.copy(), nn, isign)
        got = x.copy()
        nb.four1(got, nn, isign, L)
        assert rel(got, ref) <= tol(nn), (nn, isign, rel(got, ref))
    # round trip: inverse(forward(x)) = nn * x   (FFT_1.rs:246-267)
    y = x.copy()
    nb.four1(y, nn, 1, L)
    nb.four1(y, nn, -1, L)
    assert rel(y / nn, x) <= tol(nn)


def check_four1_batch(L, nn, count, seed=1002, scattered=False):
    big = gen(seed, 2 * nn * count)
    arrs = [big[2 * nn * b:2 * nn * (b + 1)] for b in range(count)]
    if scattered:
        arrs = [a.copy() for a in arrs]
    refs = [O.four1(a.copy(), nn, 1) for a in arrs]
    nb.FFTProcessor(L).fft_batch(arrs, 1)
    for a, r in zip(arrs, refs):
        assert rel(a, r) <= tol(nn)


def check_fourn(L, shape, seed=1003):
    n = int(np.prod(shape))
    x = gen(seed, 2 * n)
    for isign in (1, -1):
        ref = O.fourn(x.copy(), list(shape), isign)
        got = x.copy()
        nb.fourn(got, list(shape), len(shape), isign, L)
        assert rel(got, ref) <= tol(n), (shape, isign, rel(got, ref))


def check_realft(L, n, seed=1004):
    x = gen(seed, n)
    ref = O.realft(x.copy(), n, 1)
    got = x.copy()
    nb.realft(got, n, 1, L)
    assert rel(got, ref) <= tol(n), (n, rel(got, ref))
    # inverse on an arbitrary packed spectrum, and the (n/2)*x round trip
    s = gen(seed + 100, n)
    ref_i = O.realft(s.copy(), n, -1)
    got_i = s.copy()
    nb.realft(got_i, n, -1, L)
    assert rel(got_i, ref_i) <= tol(n)
    nb.realft(got, n, -1, L)
    assert rel(got * (2.0 / n), x) <= tol(n)


def check_rlft3(L, shape, seed=1006):
    nn1, nn2, nn3 = shape
    n = nn1 * nn2 * nn3
    x = gen(seed, n).reshape(shape)
    rd, rs = O.rlft3(x.copy(), np.zeros((nn1, 2 * nn2)), 1)
    d, s = x.copy(), np.zeros((nn1, 2 * nn2))
    nb.rlft3(d, s, nn1, nn2, nn3, 1, L)
    assert rel(d, rd) <= tol(n), (shape, "data", rel(d, rd))
    assert rel(s, rs) <= tol(n), (shape, "speq", rel(s, rs))
    # inverse of an arbitrary (not Hermitian-consistent) spectrum
    xi = gen(seed + 100, n).reshape(shape)
    si = gen(seed + 200, 2 * nn1 * nn2).reshape(nn1, 2 * nn2)
    ri, _ = O.rlft3(xi.copy(), si.copy(), -1)
    gi, gs = xi.copy(), si.copy()
    nb.rlft3(gi, gs, nn1, nn2, nn3, -1, L)
    assert rel(gi, ri) <= tol(n), (shape, "inverse", rel(gi, ri))
    # round trip = (nn1*nn2*nn3/2) * x
    nb.rlft3(d, s, nn1, nn2, nn3, -1, L)
    assert rel(d * (2.0 / n), x) <= tol(n)


def check_convlv(L, n, m, seed=1004):
    a = gen(seed, n)
    r = gen(seed + 1, m) / 64.0
    for pad in (nb.NRB_PAD_LITERAL, nb.NRB_PAD_NR):
        rc, ref = O.convlv(a, r, 1, pad)
        assert rc == 0
        got = nb.convlv(a, r, 1, pad, L)
        assert rel(got, ref) <= tol(n), (n, m, pad, rel(got, ref))
    # deconvolution on a well-conditioned response
    rw = 0.5 ** np.arange(min(m, 40))
    rc, ref = O.convlv(a, rw, -1, 0)
    got = nb.convlv(a, rw, -1, 0, L)
    assert rel(got, ref) <= 1e-10, (n, m, "deconv", rel(got, ref))


def check_correl(L, n, seed=1004):
    a = gen(seed, n)
    b = gen(seed + 7, n)
    rc, ref = O.correl(a, b)
    assert rc == 0
    got = nb.correl(a, b, L)
    assert rel(got, ref) <= tol(n), (n, rel(got, ref))


# ---------------------------------------------------------------- SURVEY.md 8f "next" rows
def check_twofft(L, n, seed=1010):
    a, b = gen(seed, n), gen(seed + 1, n)
    r1, r2 = O.twofft(a, b)
    f1, f2 = np.full(2 * n + 2, np.nan), np.full(2 * n + 2, np.nan)
    nb.twofft(a, b, f1, f2, L)
    assert rel(f1, r1) <= tol(n) and rel(f2, r2) <= tol(n), (n, rel(f1, r1), rel(f2, r2))
    assert f1[1] == 0.0 and f2[1] == 0.0 and not f1[2 * n:].any() and not f2[2 * n:].any()   # FFT_2.rs:60-62, :6-7


def check_twofft_batch(L, lengths, seed=1030):
    """FFT_2.rs:258 TwoFFTProcessor::process_batch: mixed lengths in one call, each result = twofft of that pair."""
    items, refs = [], []
    for i, n in enumerate(lengths):
        a, b = gen(seed + 2 * i, n), gen(seed + 2 * i + 1, n)
        items.append((a, b, np.full(2 * n + 2, np.nan), np.full(2 * n + 2, np.nan)))
        refs.append(O.twofft(a, b))
    nb.TwoFFTProcessor(L).with_optimized(False).with_threshold(16).process_batch(items)
    for (a, b, f1, f2), (r1, r2) in zip(items, refs):
        assert rel(f1, r1) <= tol(a.size) and rel(f2, r2) <= tol(a.size), (a.size, rel(f1, r1), rel(f2, r2))
    # the single-pair entry point gives the same bits as the batched one
    a, b, f1, f2 = items[0]
    g1, g2 = np.empty_like(f1), np.empty_like(f2)
    nb.TwoFFTProcessor(L).process(a, b, g1, g2)
    assert np.array_equal(f1, g1) and np.array_equal(f2, g2)
    re, im = nb.extract_real_imag(f1)
    assert np.array_equal(nb.combine_real_imag(re, im), f1)


def check_correl_normalized(L, n, seed=1011, fast=False):
    a = gen(seed, n) + 0.75
    b = 2.0 * gen(seed + 1, n) - 0.25
    rc, ref = O.correl_normalized(a, b, fast)
    assert rc == 0
    got = (nb.correl_normalized_fast if fast else nb.correl_normalized)(a, b, L)
    assert rel(got, ref) <= tol(n), (n, fast, rel(got, ref))


def check_autocorrel_fast(L, n, seed=1012):
    a = gen(seed, n)
    rc, ref = O.autocorrel_fast(a)
    assert rc == 0
    got = nb.autocorrel_fast(a, L)
    assert rel(got, ref) <= tol(n), (n, rel(got, ref))


def check_spectrum(L, npoints, seed=1013):
    c = gen(seed, 2 * npoints)
    assert rel(nb.power_spectrum(c, L), O.power_spectrum(c)) <= 1e-15
    assert rel(nb.magnitude_spectrum(c, L), O.power_spectrum(c, True)) <= 1e-15


def check_cosft1(L, n, seed=1014):
    y = np.concatenate([[123.0], gen(seed, n + 1)])          # y[0] is unused and must stay untouched
    ref = O.cosft1(y.copy(), n)
    got = y.copy()
    nb.cosft1(got, n, L)
    assert got[0] == 123.0 and rel(got[1:], ref[1:]) <= tol(n), (n, rel(got[1:], ref[1:]))


def check_cosft2(L, n, seed=1015):
    y = np.concatenate([[123.0], gen(seed, n)])
    for isign in (1, -1):
        rc, ref = O.cosft2(y.copy(), n, isign)
        got = y.copy()
        nb.cosft2(got, n, isign, L)
        assert rc == 0 and got[0] == 123.0 and rel(got[1:], ref[1:]) <= tol(n), (n, isign, rel(got[1:], ref[1:]))
    # round trip = (n/2) x   (Cos_FT2.rs:248-264 with the true factor)
    z = y.copy()
    nb.cosft2(z, n, 1, L)
    nb.cosft2(z, n, -1, L)
    assert rel(z[1:] * (2.0 / n), y[1:]) <= tol(n)


def check_sinft(L, n, seed=1016):
    y = np.concatenate([[123.0], gen(seed, n)])
    ref = O.sinft(y.copy(), n)
    got = y.copy()
    nb.sinft(got, n, L)
    assert got[0] == 123.0 and rel(got[1:], ref[1:]) <= tol(n), (n, rel(got[1:], ref[1:]))
    # the sine transform is its own inverse up to 2/n (y[1] is defined as 0)
    nb.sinft(got, n, L)
    assert rel(got[2:] * (2.0 / n), y[2:]) <= tol(n)


def check_device_resident_chain(L, shape=(8, 16, 8), seed=1017):
    """N1: upload once, rlft3 -> spectrum product -> rlft3^-1 on the device (repeated), download once."""
    from numrs_b200.device import DeviceArray, Rlft3Convolver
    n = int(np.prod(shape))
    x = gen(seed, n).reshape(shape)
    k = gen(seed + 1, n).reshape(shape) / 8.0
    conv = Rlft3Convolver(L, k)
    with DeviceArray.from_host(L, x) as xd:
        conv.apply(xd)
        conv.apply(xd)                      # a second pass without leaving the device
        got = xd.to_host().reshape(shape)
    conv.close()
    fk = np.fft.rfftn(k)
    ref = np.fft.irfftn(np.fft.rfftn(x) * fk * fk, shape, axes=(0, 1, 2))
    assert rel(got, ref) <= tol(n), rel(got, ref)
    # against the oracle's rlft3 as well (same chain on the CPU)
    d, s = x.copy(), np.zeros((shape[0], 2 * shape[1]))
    kd, ks = k.copy(), np.zeros((shape[0], 2 * shape[1]))
    O.rlft3(kd, ks, 1)
    for _ in range(2):
        O.rlft3(d, s, 1)
        dz = (d.reshape(-1)[0::2] + 1j * d.reshape(-1)[1::2]) * (kd.reshape(-1)[0::2] + 1j * kd.reshape(-1)[1::2]) * (2.0 / n)
        sz = (s.reshape(-1)[0::2] + 1j * s.reshape(-1)[1::2]) * (ks.reshape(-1)[0::2] + 1j * ks.reshape(-1)[1::2]) * (2.0 / n)
        d.reshape(-1)[0::2], d.reshape(-1)[1::2] = dz.real, dz.imag
        s.reshape(-1)[0::2], s.reshape(-1)[1::2] = sz.real, sz.imag
        O.rlft3(d, s, -1)
    assert rel(got, d) <= tol(n), rel(got, d)


def _plan_run(L, kind, n, cnt, io, aux=None, out_count=0, isign=1):
    """Device-resident plan call on `cnt` batched lines; returns (io after, out, launches)."""
    from numrs_b200.device import DeviceArray
    plan = L.plan_create(kind, [n], batch=cnt)
    launches = plan.num_launches(isign)
    d_io = DeviceArray.from_host(L, io)
    d_aux = DeviceArray.from_host(L, aux) if aux is not None else None
    d_out = DeviceArray(L, out_count) if out_count else None
    plan.exec(d_io.ptr, d_aux.ptr if d_aux else 0, d_out.ptr if d_out else 0, isign=isign)
    got_io = d_io.to_host()
    got_out = d_out.to_host() if d_out else None
    for d in (d_io, d_aux, d_out):
        if d is not None:
            d.free()
    plan.destroy()
    return got_io, got_out, launches


def check_trig_batch(L, n, cnt, seed=1040):
    """cosft1 / cosft2 (both signs) / sinft on `cnt` batched lines of the reference's 1-based arrays through the plan API:
    the one-kernel path (trig_fused.cuh) against the oracle line by line and against the multi-launch path; element 0 of
    every line (unused by the reference) must stay untouched.  cnt lines of odd length put the lines on alternating
    16-byte alignments."""
    routines = ((nb.KIND_COSFT1, n + 2, 1, lambda y: O.cosft1(y, n)),
                (nb.KIND_COSFT2, n + 1, 1, lambda y: O.cosft2(y, n, 1)[1]),
                (nb.KIND_COSFT2, n + 1, -1, lambda y: O.cosft2(y, n, -1)[1]),
                (nb.KIND_SINFT, n + 1, 1, lambda y: O.sinft(y, n)))
    for kind, ld, isign, ref_fn in routines:
        y = gen(seed + kind, cnt * ld)
        y[0::ld] = 123.0 + np.arange(cnt)
        L.set_option("trig_fused", 1)
        got, _, launches = _plan_run(L, kind, n, cnt, y, isign=isign)
        assert launches == 1, (kind, n, launches)
        L.set_option("trig_fused", 0)
        old, _, launches0 = _plan_run(L, kind, n, cnt, y, isign=isign)
        L.set_option("trig_fused", 1)
        assert launches0 > 1
        for i in range(cnt):
            ref = ref_fn(y[i * ld:(i + 1) * ld].copy())
            nd = n + 1 if kind == nb.KIND_COSFT1 else n
            g = got[i * ld:(i + 1) * ld]
            assert g[0] == y[i * ld], (kind, n, i)
            assert rel(g[1:1 + nd], ref[1:1 + nd]) <= tol(n), (kind, isign, n, i, rel(g[1:1 + nd], ref[1:1 + nd]))
            assert rel(g[1:1 + nd], old[i * ld + 1:i * ld + 1 + nd]) <= 1e-13, (kind, isign, n, i)
            if ld > nd + 1:                                   # cosft1's line has no tail; keep the check generic
                assert np.array_equal(g[1 + nd:], y[i * ld + 1 + nd:(i + 1) * ld])


def check_twofft_plan(L, n, cnt, seed=1050):
    """twofft on `cnt` batched signal pairs through the plan API: one-kernel path against the oracle and the three-launch path."""
    a, b = gen(seed, n * cnt) + 0.25, gen(seed + 1, n * cnt)
    per = 2 * n + 2
    res = {}
    for flag in (1, 0):
        L.set_option("trig_fused", flag)
        _, f, launches = _plan_run(L, nb.KIND_TWOFFT, n, cnt, a, aux=b, out_count=2 * cnt * per)
        assert (launches == 1) == (flag == 1), (n, flag, launches)
        res[flag] = f
    L.set_option("trig_fused", 1)
    for i in range(cnt):
        r1, r2 = O.twofft(a[i * n:(i + 1) * n], b[i * n:(i + 1) * n])
        assert rel(res[1][i * per:(i + 1) * per], r1) <= tol(n), (n, i)
        assert rel(res[1][(cnt + i) * per:(cnt + i + 1) * per], r2) <= tol(n), (n, i)
    assert rel(res[1], res[0]) <= 1e-14
