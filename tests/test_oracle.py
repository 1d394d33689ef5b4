"""CPU tier: pins the oracle (oracle/nr_oracle.c) against numpy, a 50-digit mpmath DFT, the
committed golden vectors and the known answers held by the reference's own unit tests."""
import json
import os

import numpy as np
import pytest

import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden_small.npz"))
KNOWN = json.load(open(os.path.join(HERE, "golden", "reference_known_answers.json")))


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def c(z):
    return z[0::2] + 1j * z[1::2]


def test_generator_c_equals_numpy():
    assert np.array_equal(O.fill_uniform(1001, 12345, 4096), O.fill_uniform_py(1001, 12345, 4096))
    x = O.fill_uniform(1006, 0, 1 << 16)
    assert x.min() >= -1.0 and x.max() < 1.0 and abs(x.mean()) < 0.02


@pytest.mark.parametrize("nn", [1, 2, 4, 8, 256, 1024, 4096, 1 << 15])
def test_four1_vs_numpy(nn):
    x = O.fill_uniform(1001, 0, 2 * nn)
    for isign, ref in ((1, np.fft.ifft(c(x)) * nn), (-1, np.fft.fft(c(x)))):
        for fn in (O.four1, O.four1_optimized):
            assert rel(c(fn(x.copy(), nn, isign)), ref) < 1e-13 * max(1, np.log2(max(nn, 2)))
    assert np.allclose(O.four1(x.copy(), nn, 1, mt=True), O.four1(x.copy(), nn, 1), rtol=0, atol=0)


def test_four1_vs_mpmath():
    import mpmath as mp
    mp.mp.dps = 50
    nn = 32
    x = O.fill_uniform(1001, 7, 2 * nn)
    z = [mp.mpc(x[2 * k], x[2 * k + 1]) for k in range(nn)]
    ref = [sum(z[j] * mp.e ** (2j * mp.pi * j * k / nn) for j in range(nn)) for k in range(nn)]
    got = c(O.four1(x.copy(), nn, 1))
    err = max(abs(complex(r) - g) for r, g in zip(ref, got))
    assert err < 1e-13


# ---- every routine the reference's own tests do not pin is checked against a 50-digit DFT at N <= 64 ----
def _mp_dft(vals, shape, sign):
    """Unnormalised DFT of the (possibly N-D) array `vals` (flat, row-major) with exp(sign * 2 pi i sum k_d j_d / n_d), 50 digits."""
    import itertools

    import mpmath as mp
    mp.mp.dps = 50
    idx = list(itertools.product(*[range(n) for n in shape]))
    roots = [[mp.e ** (sign * 2j * mp.pi * m / n) for m in range(n)] for n in shape]
    out = []
    for k in idx:
        acc = mp.mpc(0)
        for j, v in zip(idx, vals):
            w = mp.mpc(1)
            for d, n in enumerate(shape):
                w *= roots[d][(k[d] * j[d]) % n]
            acc += v * w
        out.append(acc)
    return out


@pytest.mark.parametrize("n", [2, 4, 16, 64])
def test_realft_vs_mpmath(n):
    """NR realft (Real_FT.rs:4-21, ledger D1/D2): forward = packed half spectrum of F_k = sum x_j e^{+2 pi i jk/n}; the
    inverse of that spectrum returns (n/2) x."""
    import mpmath as mp
    x = O.fill_uniform(1010, 3, n)
    F = _mp_dft([mp.mpf(v) for v in x], (n,), +1)
    got = O.realft(x.copy(), n, 1)
    want = [F[0].real, F[n // 2].real] + [p for k in range(1, n // 2) for p in (F[k].real, F[k].imag)]
    assert max(abs(float(w) - g) for w, g in zip(want, got)) < 1e-13 * max(1.0, float(np.linalg.norm(x)))
    back = O.realft(got.copy(), n, -1)
    assert np.max(np.abs(back / (n / 2) - x)) < 1e-14 * n


@pytest.mark.parametrize("shape", [(4, 8), (2, 4, 8), (8, 8), (4, 4, 4)])
def test_fourn_vs_mpmath(shape):
    """NR in-memory fourn (call shape Real_FT3.rs:35, ledger D3): last index fastest, exp(isign * 2 pi i sum k_d j_d / nn_d)."""
    import mpmath as mp
    n = int(np.prod(shape))
    x = O.fill_uniform(1011, 5, 2 * n)
    z = [mp.mpc(x[2 * k], x[2 * k + 1]) for k in range(n)]
    for isign in (1, -1):
        ref = _mp_dft(z, shape, isign)
        got = c(O.fourn(x.copy(), list(shape), isign))
        assert max(abs(complex(r) - g) for r, g in zip(ref, got)) < 1e-13 * np.sqrt(n)


@pytest.mark.parametrize("shape", [(2, 2, 4), (4, 2, 8), (2, 4, 4), (4, 4, 4)])
def test_rlft3_vs_mpmath(shape):
    """NR rlft3 (Real_FT3.rs:8-141, ledger D3/D4): data[i1][i2][2k..2k+1] = F(i1, i2, k) for k < nn3/2 and
    speq[i1][2 i2 .. 2 i2 + 1] = F(i1, i2, nn3/2) with F = sum x e^{+2 pi i (...)}; the inverse returns (N/2) x."""
    import mpmath as mp
    nn1, nn2, nn3 = shape
    n = nn1 * nn2 * nn3
    x = O.fill_uniform(1012, 9, n).reshape(shape)
    F = _mp_dft([mp.mpf(v) for v in x.ravel()], shape, +1)
    F = np.array([complex(v) for v in F]).reshape(shape)
    d, s = O.rlft3(x.copy(), np.zeros((nn1, 2 * nn2)), 1)
    dc = d.reshape(nn1, nn2, nn3 // 2, 2)
    assert np.max(np.abs((dc[..., 0] + 1j * dc[..., 1]) - F[:, :, :nn3 // 2])) < 1e-13 * np.sqrt(n)
    sc = s.reshape(nn1, nn2, 2)
    assert np.max(np.abs((sc[..., 0] + 1j * sc[..., 1]) - F[:, :, nn3 // 2])) < 1e-13 * np.sqrt(n)
    back, _ = O.rlft3(d.copy(), s.copy(), -1)
    assert np.max(np.abs(back * (2.0 / n) - x)) < 1e-14 * n


def test_four1_reference_round_trip():
    # FFT_1.rs:246-267
    n = 1024
    t = np.arange(n) / n
    sig = np.sin(2 * np.pi * 5 * t) + 0.5 * np.cos(2 * np.pi * 20 * t)
    z = np.zeros(2 * n)
    z[0::2] = sig
    O.four1(z, n, 1)
    O.four1(z, n, -1)
    assert np.max(np.abs(z[0::2] / n - sig)) < KNOWN["four1_round_trip"]["abs_tol"]


@pytest.mark.parametrize("shape", [(4, 8, 2), (8, 16), (16,), (2, 2), (32, 4, 8), (64, 64)])
def test_fourn_vs_numpy(shape):
    n = int(np.prod(shape))
    x = O.fill_uniform(1003, 0, 2 * n)
    for isign in (1, -1):
        ref = np.fft.ifftn(c(x).reshape(shape)) * n if isign == 1 else np.fft.fftn(c(x).reshape(shape))
        assert rel(c(O.fourn(x.copy(), list(shape), isign)), ref.ravel()) < 1e-13 * np.log2(n)
        assert np.array_equal(O.fourn(x.copy(), list(shape), isign, mt=True), O.fourn(x.copy(), list(shape), isign))


def test_fourn_validation_rules():
    for case in KNOWN["fourn_validation"]["cases"]:
        rc = O.fourn_validate(case["nn"], case["ndim"], case["isign"])
        assert (rc == 0) == case["ok"]


@pytest.mark.parametrize("n", [2, 4, 8, 16, 256, 4096, 1 << 14])
def test_realft_vs_numpy(n):
    x = O.fill_uniform(1004, 0, n)
    y = O.realft(x.copy(), n, 1)
    F = np.conj(np.fft.rfft(x))
    assert abs(y[0] - F[0].real) < 1e-12 and abs(y[1] - F[n // 2].real) < 1e-12
    if n > 2:
        assert rel(y[2::2] + 1j * y[3::2], F[1:n // 2]) < 1e-13 * np.log2(n)
    assert rel(O.realft(y.copy(), n, -1) * 2 / n, x) < 1e-13 * np.log2(n)   # round trip = (n/2) x


@pytest.mark.parametrize("shp", [(8, 8, 8), (4, 16, 8), (1, 4, 4), (2, 2, 2), (16, 8, 32)])
def test_rlft3_vs_numpy(shp):
    x = O.fill_uniform(1006, 0, int(np.prod(shp))).reshape(shp)
    d, s = O.rlft3(x.copy(), np.zeros((shp[0], 2 * shp[1])), 1)
    F = np.conj(np.fft.rfftn(x))
    dd = d.reshape(shp[0], shp[1], shp[2] // 2, 2)
    ss = s.reshape(shp[0], shp[1], 2)
    assert rel(dd[..., 0] + 1j * dd[..., 1], F[..., :shp[2] // 2]) < 1e-13 * np.log2(x.size)
    assert rel(ss[..., 0] + 1j * ss[..., 1], F[..., shp[2] // 2]) < 1e-13 * np.log2(x.size)
    d2, _ = O.rlft3(d.copy(), s.copy(), -1)
    assert rel(d2 * 2 / x.size, x) < 1e-13 * np.log2(x.size)               # round trip = N/2 * x


def test_rlft3_reference_ramp_round_trip():
    # Real_FT3.rs:268-311 uses the ramp (i+j+k) on 8^3; the true factor is N/2 (ledger D8)
    i, j, k = np.meshgrid(np.arange(8), np.arange(8), np.arange(8), indexing="ij")
    x = (i + j + k).astype(np.float64)
    d, s = O.rlft3(x.copy(), np.zeros((8, 16)), 1)
    d, s = O.rlft3(d, s, -1)
    assert np.allclose(d / (8 * 8 * 8 / 2), x, atol=1e-10)


def test_convlv_known_answers_and_errors():
    ka = KNOWN["convlv_basic"]
    rc, y = O.convlv(ka["data"], ka["respns"], ka["isign"])
    assert rc == 0
    for idx, val in ka["expect_at"].items():
        assert abs(y[int(idx)] - val) < ka["abs_tol"]
    codes = {"EmptyInput": -1, "ResponseTooLong": -2, "InvalidIsign": -3}
    for case in KNOWN["convlv_errors"]["cases"]:
        rc, _ = O.convlv(case["data"], case["respns"], case["isign"])
        assert rc == codes[case["err"]]


@pytest.mark.parametrize("n,m", [(64, 5), (256, 2), (1024, 33), (4096, 4096)])
def test_convlv_vs_numpy(n, m):
    a = O.fill_uniform(1004, 0, n)
    r = O.fill_uniform(1005, 0, m) / 64
    for pad in (0, 1):
        rc, y = O.convlv(a, r, 1, pad)
        p = O.pad_response(r, n, pad)
        assert rc == 0 and rel(y, np.fft.irfft(np.fft.rfft(a) * np.fft.rfft(p), n)) < 1e-13 * np.log2(n)
    # NR padding really is NR's wrap-around
    p = O.pad_response(np.arange(1.0, 6.0), 16, 1)
    assert list(p[:3]) == [1, 2, 3] and list(p[-2:]) == [4, 5] and not p[3:-2].any()
    # literal padding, m = 2 and m = 3 (SURVEY.md ledger L1)
    assert list(O.pad_response([7.0, 9.0], 4, 0)) == [7, 9, 0, 0]
    assert list(O.pad_response([1.0, 2.0, 3.0], 6, 0)) == [3, 2, 3, 0, 0, 1]


def test_correl_known_answers_and_errors():
    rc, y = O.correl(KNOWN["correl_basic"]["a"], KNOWN["correl_basic"]["b"])
    assert rc == 0 and y[0] == 30.0 and y[0] > y[1]
    rc, y = O.correl(*[KNOWN["correl_direct_small"][k] for k in ("a", "b")])
    assert list(y) == KNOWN["correl_direct_small"]["expect"]
    rc, y = O.correl(KNOWN["autocorrel"]["a"], KNOWN["autocorrel"]["a"])
    assert y[0] == 10.0
    for (a, b), e in zip(KNOWN["correl_batch"]["pairs"], KNOWN["correl_batch"]["expect0"]):
        assert O.correl(a, b)[1][0] == e
    codes = {"EmptyInput": -1, "LengthMismatch": -4}
    for case in KNOWN["correl_errors"]["cases"]:
        assert O.correl(case["a"], case["b"])[0] == codes[case["err"]]


@pytest.mark.parametrize("n", [64, 1024, 1 << 14])
def test_correl_vs_numpy(n):
    a = O.fill_uniform(1004, 0, n)
    b = O.fill_uniform(1011, 0, n)
    rc, y = O.correl(a, b)
    assert rc == 0 and rel(y, np.fft.irfft(np.fft.rfft(a) * np.conj(np.fft.rfft(b)), n)) < 1e-13 * np.log2(n)


def test_oracle_matches_golden_fixtures():
    for nn in (8, 64, 1024):
        for s, t in ((1, "p"), (-1, "m")):
            assert np.array_equal(O.four1(G[f"four1_{nn}_in"].copy(), nn, s), G[f"four1_{nn}_{t}"])
    for shape in ((4, 8, 2), (8, 16), (16, 4, 8)):
        tag = "x".join(map(str, shape))
        for s, t in ((1, "p"), (-1, "m")):
            assert np.array_equal(O.fourn(G[f"fourn_{tag}_in"].copy(), list(shape), s), G[f"fourn_{tag}_{t}"])
    for n in (8, 256, 2048):
        assert np.array_equal(O.realft(G[f"realft_{n}_in"].copy(), n, 1), G[f"realft_{n}_fwd"])
    for shp in ((8, 8, 8), (4, 16, 8), (2, 4, 32)):
        tag = "x".join(map(str, shp))
        d, s = O.rlft3(G[f"rlft3_{tag}_in"].copy(), np.zeros((shp[0], 2 * shp[1])), 1)
        assert np.array_equal(d, G[f"rlft3_{tag}_data"]) and np.array_equal(s, G[f"rlft3_{tag}_speq"])
    assert np.array_equal(O.convlv(G["convlv_128_9_in"], G["convlv_128_9_resp"], 1)[1], G["convlv_128_9_out"])
    assert np.array_equal(O.correl(G["correl_128_a"], G["correl_128_b"])[1], G["correl_128_out"])


# ---------------------------------------------------------------- SURVEY.md 8f "next" rows (N2, N4)
@pytest.mark.parametrize("n", [1, 2, 4, 16, 256, 4096])
def test_twofft_vs_numpy(n):
    """NR twofft (ledger D9): both spectra equal the e^{+} DFT of each real signal (numpy: n * ifft)."""
    a, b = O.fill_uniform(1010, 0, n), O.fill_uniform(1011, 0, n)
    f1, f2 = O.twofft(a, b)
    assert rel(c(f1[:2 * n]), n * np.fft.ifft(a)) < 1e-13 * max(1, np.log2(max(n, 2)))
    assert rel(c(f2[:2 * n]), n * np.fft.ifft(b)) < 1e-13 * max(1, np.log2(max(n, 2)))
    assert not f1[2 * n:].any() and not f2[2 * n:].any() and f1[1] == 0.0 and f2[1] == 0.0


@pytest.mark.parametrize("n", [2, 5, 32, 64, 1024])
def test_correl_normalized_vs_numpy(n):
    a = O.fill_uniform(1011, 0, n) + 0.75
    b = 2.0 * O.fill_uniform(1012, 0, n) - 0.25
    an, bn = (a - a.mean()) / a.std(), (b - b.mean()) / b.std()
    for fast in (False, True):
        rc, got = O.correl_normalized(a, b, fast)
        assert rc == 0
        if n <= 32:     # Correlation.rs:19-21 / :251-263 direct lags; only the fast variant divides by n
            ref = np.array([np.dot(an[l:], bn[:n - l]) for l in range(n)]) * (1.0 / n if fast else 1.0)
        else:
            ref = np.fft.irfft(np.fft.rfft(an) * np.conj(np.fft.rfft(bn)), n)
        assert rel(got, ref) < 1e-13 * max(1, np.log2(n))
    assert O.correl_normalized(np.ones(8), np.arange(8.0))[0] == -8        # CorrelError::ZeroStdDev
    assert O.correl_normalized([], [1.0])[0] == -1 and O.correl_normalized([1.0, 2.0], [1.0])[0] == -4
    # Correlation.rs:505-512
    assert abs(O.correl_normalized([1.0, 2.0, 3.0, 4.0], [1.0, 2.0, 3.0, 4.0], True)[1][0] - 1.0) < 1e-10


@pytest.mark.parametrize("n", [3, 32, 64, 4096])
def test_autocorrel_fast_and_spectra_vs_numpy(n):
    a = O.fill_uniform(1012, 0, n)
    rc, got = O.autocorrel_fast(a)
    ref = (np.array([np.dot(a[l:], a[:n - l]) for l in range(n)]) if n <= 32
           else np.fft.irfft(np.abs(np.fft.rfft(a)) ** 2, n))
    assert rc == 0 and rel(got, ref) < 1e-13 * max(1, np.log2(n))
    z = O.fill_uniform(1013, 0, 2 * n)
    assert np.array_equal(O.power_spectrum(z), z[0::2] ** 2 + z[1::2] ** 2)
    assert np.allclose(O.power_spectrum(z, True), np.abs(c(z)), rtol=1e-15, atol=0)


@pytest.mark.parametrize("n", [2, 4, 8, 64, 1024, 1 << 14])
def test_cosft_sinft_vs_scipy(n):
    """N3 (ledger D11): NR cosft1 = DCT-I / 2, cosft2(+1) = DCT-II / 2, cosft2(-1) its inverse times n/2,
    sinft = DST-I / 2 (scipy's unnormalised definitions carry a factor 2)."""
    from scipy.fft import dct, dst
    lim = 1e-12 * max(1, np.log2(n))
    f = O.fill_uniform(1014, 0, n + 1)
    y = np.concatenate([[7.0], f])
    O.cosft1(y, n)
    assert y[0] == 7.0 and rel(y[1:], dct(f, type=1) / 2) < lim
    f = O.fill_uniform(1015, 0, n)
    y = np.concatenate([[7.0], f])
    assert O.cosft2(y, n, 1)[0] == 0 and rel(y[1:], dct(f, type=2) / 2) < lim
    assert O.cosft2(y, n, -1)[0] == 0 and rel(y[1:] * (2.0 / n), f) < lim
    assert O.cosft2(y, n, 0)[0] == -3                      # Cos_FT2.rs:11 / :266-272
    y = np.concatenate([[7.0], f])
    O.sinft(y, n)
    assert y[1] == 0.0 and rel(y[2:], dst(f[1:], type=1) / 2) < lim
