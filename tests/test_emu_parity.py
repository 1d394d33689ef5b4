"""CPU tier: planner + kernel index arithmetic, validated by running the device source under
the TEST-ONLY host emulation (tests/emu) through the same C ABI, against the oracle.
This does not measure or ship anything; GPU parity proper is tests/test_gpu_parity.py."""
import json
import os

import numpy as np
import pytest

import cases
import numrs_b200 as nb
import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden_small.npz"))
KNOWN = json.load(open(os.path.join(HERE, "golden", "reference_known_answers.json")))


@pytest.mark.parametrize("nn", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_four1_single_pass(emu, nn):
    cases.check_four1(emu, nn)


@pytest.mark.parametrize("nn", [1 << 14, 1 << 16])
def test_four1_four_step(emu, nn):
    cases.check_four1(emu, nn)


@pytest.mark.parametrize("nn", [64, 512, 4096, 1 << 13])
def test_four1_multi_step_three_factors(emu, nn):
    emu.set_option("row_max_log2", 4)
    emu.set_option("col_max_log2", 3)
    cases.check_four1(emu, nn)


def test_four1_identity_sizes_and_errors(emu):
    x = np.array([3.0, 4.0])
    nb.four1(x, 1, 1, emu)                      # nn = 1 is the identity (FFT_1.rs:5-44)
    assert list(x) == [3.0, 4.0]
    assert emu.four1(np.zeros(12), 6, 1) == nb._lib.NRB_ERR_NOT_POW2
    assert emu.four1(np.zeros(16), 8, 0) == nb._lib.NRB_ERR_INVALID_ISIGN


def test_four1_batch(emu):
    cases.check_four1_batch(emu, 256, 37)                   # contiguous slices, ragged tile
    cases.check_four1_batch(emu, 64, 5, scattered=True)     # separate allocations
    # mixed lengths in one call (FFT_1.rs:185-189 allows any slice lengths)
    arrs = [O.fill_uniform(5, 0, 2 * n) for n in (8, 64, 8, 1024, 1)]
    refs = [O.four1(a.copy(), a.size // 2, -1) for a in arrs]
    nb.FFTProcessor(emu).fft_batch(arrs, -1)
    for a, r in zip(arrs, refs):
        assert cases.rel(a, r) <= cases.tol(a.size)


@pytest.mark.parametrize("shape", [(2,), (4, 8, 2), (8, 16), (16,), (2, 2), (32, 4, 8), (64, 64), (8, 8, 8, 4),
                                   (2, 1024), (1024, 2), (128, 32), (2048, 4), (4, 4096)])
def test_fourn(emu, shape):
    cases.check_fourn(emu, shape)


@pytest.mark.parametrize("shape", [(64, 64), (16, 32, 8), (256, 4)])
def test_fourn_multi_step_axes(emu, shape):
    emu.set_option("row_max_log2", 3)
    emu.set_option("col_max_log2", 2)
    cases.check_fourn(emu, shape)


def test_fourn_validation(emu):
    for case in KNOWN["fourn_validation"]["cases"]:
        n = int(np.prod(case["nn"]))
        if case["ok"]:
            nb.fourn(np.zeros(2 * n), case["nn"], case["ndim"], case["isign"], emu)
        else:
            with pytest.raises(ValueError):
                nb.fourn(np.zeros(2 * n), case["nn"], case["ndim"], case["isign"], emu)
    with pytest.raises(ValueError):
        nb.fourn(np.zeros(4), [2], 0, 1, emu)       # ndim == 0
    with pytest.raises(ValueError):
        nb.fourn(np.zeros(4), [2], 2, 1, emu)       # ndim > nn.len()


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 256, 1024, 4096, 1 << 14, 1 << 15])
def test_realft(emu, n):
    cases.check_realft(emu, n)


@pytest.mark.parametrize("n", [16, 256, 4096])
def test_realft_large_line_path(emu, n):
    emu.set_option("row_max_log2", 2)       # forces c2c passes + standalone untangle kernel
    emu.set_option("col_max_log2", 3)
    cases.check_realft(emu, n)


def test_realft_asserts_and_batch(emu):
    with pytest.raises(AssertionError):
        nb.realft(np.zeros(6), 5, 1, emu)            # Real_FT.rs:5
    with pytest.raises(AssertionError):
        nb.realft(np.zeros(4), 8, 1, emu)            # Real_FT.rs:6
    xs = [O.fill_uniform(9, i * 512, 512) for i in range(5)]
    refs = [O.realft(x.copy(), 512, 1) for x in xs]
    nb.RealFTProcessor(emu).process_batch([(x, 512, 1) for x in xs])
    for x, r in zip(xs, refs):
        assert cases.rel(x, r) <= cases.tol(512)
    # isign other than 1 means inverse (Real_FT.rs:10,15)
    a, b = refs[0].copy(), refs[0].copy()
    nb.realft(a, 512, -1, emu)
    nb.realft(b, 512, 7, emu)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("shp", [(8, 8, 8), (4, 16, 8), (1, 4, 4), (2, 2, 2), (16, 8, 32), (1, 1, 2), (2, 1, 4),
                                 (32, 32, 32), (4, 4, 2), (8, 64, 16), (1, 1, 64), (64, 1, 2)])
def test_rlft3(emu, shp):
    cases.check_rlft3(emu, shp)


def test_rlft3_grouped_and_multi_step(emu):
    emu.set_option("l2_group_bytes", 4096)   # several x-plane groups
    cases.check_rlft3(emu, (16, 16, 16))
    emu.set_option("col_max_log2", 2)
    emu.set_option("row_max_log2", 2)
    cases.check_rlft3(emu, (16, 32, 32))


@pytest.mark.parametrize("shp,lag", [((4, 256, 256), 16), ((6 - 2, 256, 512), 1), ((8, 512, 256), 3)])
def test_rlft3_fused_zy_program(emu, shp, lag):
    """Shapes that take the fused z+y launch (under emulation the pair runs back to back)."""
    emu.set_option("fuse_zy", 1)
    emu.set_option("fuse_lag", lag)
    cases.check_rlft3(emu, shp)
    plan = emu.plan_create(nb.KIND_RLFT3, list(shp))
    # fused pair, speq y, x, speq x (5 when the build's ROW and COL CTA sizes differ: no fused kernels)
    assert plan.num_launches(1) in (4, 5) and plan.num_launches(-1) == plan.num_launches(1)
    plan.destroy()
    emu.set_option("fuse_zy", 0)
    plan = emu.plan_create(nb.KIND_RLFT3, list(shp))
    assert plan.num_launches(1) == 5
    plan.destroy()


def test_rlft3_asserts(emu):
    d, s = np.zeros((4, 4, 4)), np.zeros((4, 8))
    with pytest.raises(AssertionError):
        nb.rlft3(d, s, 4, 4, 4, 0, emu)              # Real_FT3.rs:17
    with pytest.raises(AssertionError):
        nb.rlft3(d, s, 4, 4, 8, 1, emu)              # Real_FT3.rs:18
    with pytest.raises(AssertionError):
        nb.rlft3(d, np.zeros((4, 4)), 4, 4, 4, 1, emu)   # Real_FT3.rs:19


@pytest.mark.parametrize("n,m", [(4, 2), (2, 1), (2, 2), (64, 5), (64, 64), (256, 9), (1024, 33), (1 << 14, 100),
                                 (1 << 15, 4096)])
def test_convlv(emu, n, m):
    cases.check_convlv(emu, n, m)


@pytest.mark.parametrize("n,m", [(16, 3), (64, 5), (1024, 33), (4096, 100)])
def test_convlv_correl_fused_spectral_path(emu, n, m):
    """Long-line path: c2c passes + one fused untangle/multiply/untangle kernel (AUX_SPECTRAL_Z)."""
    emu.set_option("row_max_log2", 2)
    emu.set_option("col_max_log2", 3)
    cases.check_convlv(emu, n, m)
    if n > 32:
        cases.check_correl(emu, n)


@pytest.mark.parametrize("nn", [1 << 18, 1 << 20])
def test_four1_four_step_small_line_counts(emu, nn):
    """XPOSE tiles with 8 and 4 lines (XOR-swizzled shared memory for L < 8)."""
    cases.check_four1(emu, nn)


def test_convlv_reference_known_answers(emu):
    ka = KNOWN["convlv_basic"]
    y = nb.convlv(ka["data"], ka["respns"], ka["isign"], _L=emu)
    for idx, val in ka["expect_at"].items():
        assert abs(y[int(idx)] - val) < ka["abs_tol"]
    assert abs(nb.ConvlvProcessor(emu).process(ka["data"], ka["respns"], 1)[1] - 3.0) < 1e-10   # Convolve.rs:445-454
    for case in KNOWN["convlv_errors"]["cases"]:
        with pytest.raises(nb.ConvlvError) as ei:
            nb.convlv(case["data"], case["respns"], case["isign"], _L=emu)
        assert ei.value.kind == case["err"]
    with pytest.raises(nb.ConvlvError) as ei:      # odd n: the reference panics in realft (Convolve.rs:427-442)
        nb.convlv([1.0, 2.0, 3.0], [1.0, 1.0], 1, _L=emu)
    assert ei.value.kind == nb.ConvlvError.FftError


def test_convlv_batch(emu):
    sigs = [O.fill_uniform(1004, i * 1024, 1024) for i in range(7)]
    r = O.fill_uniform(1005, 0, 17) / 64
    emu.set_option("batch_group_bytes", 3 * 1024 * 8)   # several signal groups
    outs = nb.convlv_batch(sigs, r, 1, _L=emu)
    for s, o in zip(sigs, outs):
        assert cases.rel(o, O.convlv(s, r, 1)[1]) <= cases.tol(1024)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 32, 64, 128, 1024, 1 << 14])
def test_correl(emu, n):
    cases.check_correl(emu, n)


def test_correl_reference_known_answers(emu):
    y = nb.correl(KNOWN["correl_basic"]["a"], KNOWN["correl_basic"]["b"], emu)
    assert y[0] == 30.0 and y[0] > y[1]
    assert list(nb.correl([1.0, 2.0], [1.0, 2.0], emu)) == KNOWN["correl_direct_small"]["expect"]
    assert nb.autocorrel([1.0, 2.0, 1.0, 2.0], emu)[0] == 10.0
    outs = nb.correl_batch([(np.array(a, float), np.array(b, float)) for a, b in KNOWN["correl_batch"]["pairs"]], emu)
    assert [o[0] for o in outs] == KNOWN["correl_batch"]["expect0"]
    for case in KNOWN["correl_errors"]["cases"]:
        with pytest.raises(nb.CorrelError) as ei:
            nb.correl(case["a"], case["b"], emu)
        assert ei.value.kind == case["err"]


def test_correl_batch_large(emu):
    a = [O.fill_uniform(1, i * 256, 256) for i in range(5)]
    b = [O.fill_uniform(2, i * 256, 256) for i in range(5)]
    emu.set_option("batch_group_bytes", 2 * 256 * 8)
    outs = nb.correl_batch(list(zip(a, b)), emu)
    for x, y, o in zip(a, b, outs):
        assert cases.rel(o, O.correl(x, y)[1]) <= cases.tol(256)


def test_golden_fixtures(emu):
    for nn in (8, 64, 1024):
        for s, t in ((1, "p"), (-1, "m")):
            x = G[f"four1_{nn}_in"].copy()
            nb.four1(x, nn, s, emu)
            assert cases.rel(x, G[f"four1_{nn}_{t}"]) <= cases.tol(nn)
    for shp in ((8, 8, 8), (4, 16, 8), (2, 4, 32)):
        tag = "x".join(map(str, shp))
        d, s = G[f"rlft3_{tag}_in"].copy(), np.zeros((shp[0], 2 * shp[1]))
        nb.rlft3(d, s, *shp, 1, emu)
        assert cases.rel(d, G[f"rlft3_{tag}_data"]) <= cases.tol(d.size)
        assert cases.rel(s, G[f"rlft3_{tag}_speq"]) <= cases.tol(d.size)


def test_plan_api_device_pointers(emu):
    """Device-resident plan API (under emulation 'device' pointers are host pointers)."""
    n = 4096
    x = O.fill_uniform(3, 0, 2 * n * 3)
    ref = np.concatenate([O.four1(x[2 * n * b:2 * n * (b + 1)].copy(), n, 1) for b in range(3)])
    plan = emu.plan_create(nb.KIND_FOUR1, [n], batch=3)
    assert plan.num_launches(1) == 1 and plan.workspace_bytes() == 0
    plan.exec(x.ctypes.data, isign=1)
    assert cases.rel(x, ref) <= cases.tol(n)
    prof = plan.profile(x.ctypes.data, isign=-1)
    assert prof[0][0] == "fft_row_plain_n4096_m_L3" and prof[0][1] == 2 * 16 * n * 3
    plan.destroy()
    big = emu.plan_create(nb.KIND_FOUR1, [1 << 16], batch=1)
    assert big.num_launches(1) == 2 and big.workspace_bytes() == (1 << 16) * 16
    big.destroy()
    with pytest.raises(nb.NrbError):
        emu.plan_create(nb.KIND_FOURN, [8, 1], batch=1)


def test_device_fill_matches_generator(emu):
    out = np.zeros(5000)
    emu.fill_uniform_device(out.ctypes.data, 1006, 77, out.size)
    assert np.array_equal(out, O.fill_uniform(1006, 77, out.size))


# ---------------------------------------------------------------- SURVEY.md 8f "next" rows (N2, N4)
@pytest.mark.parametrize("n", [1, 2, 4, 16, 256, 4096, 8192, 1 << 14])
def test_twofft(emu, n):
    cases.check_twofft(emu, n)


def test_twofft_multi_step_and_reference_properties(emu):
    emu.set_option("row_max_log2", 3)       # dense workspace path
    emu.set_option("col_max_log2", 3)
    cases.check_twofft(emu, 256)
    # FFT_2.rs:392-428 test_twofft_correctness (with the 0-based mirror n - k, ledger D9)
    n = 256
    t = np.arange(n) / n
    f1, f2 = np.zeros(2 * n + 2), np.zeros(2 * n + 2)
    nb.twofft(np.sin(2 * np.pi * 5 * t), np.cos(2 * np.pi * 10 * t), f1, f2, emu)
    assert abs(f1[1]) < 1e-10 and abs(f2[1]) < 1e-10
    for k in range(1, n // 2):
        assert abs(f1[2 * k] - f1[2 * (n - k)]) < 1e-10 and abs(f1[2 * k + 1] + f1[2 * (n - k) + 1]) < 1e-10
    # the spectra themselves: sin(2 pi 5 t) -> +-i n/2 at bins 5 / n-5; cos(2 pi 10 t) -> n/2 at bins 10 / n-10
    assert abs(f1[2 * 5 + 1] - n / 2) < 1e-9 and abs(f2[2 * 10] - n / 2) < 1e-9
    with pytest.raises(AssertionError):
        nb.twofft(np.zeros(4), np.zeros(5), np.zeros(10), np.zeros(10), emu)      # FFT_2.rs:5
    with pytest.raises(AssertionError):
        nb.twofft(np.zeros(4), np.zeros(4), np.zeros(8), np.zeros(10), emu)       # FFT_2.rs:6


@pytest.mark.parametrize("n", [1, 2, 5, 31, 32, 64, 128, 4096, 1 << 15])
@pytest.mark.parametrize("fast", [False, True])
def test_correl_normalized(emu, n, fast):
    if n == 1:
        with pytest.raises(nb.CorrelError) as ei:       # a single sample has zero std
            (nb.correl_normalized_fast if fast else nb.correl_normalized)([1.0], [2.0], emu)
        assert ei.value.kind == nb.CorrelError.ZeroStdDev
        return
    cases.check_correl_normalized(emu, n, fast=fast)


def test_correl_normalized_long_line_path_and_errors(emu):
    emu.set_option("row_max_log2", 2)
    emu.set_option("col_max_log2", 3)
    cases.check_correl_normalized(emu, 1024)
    cases.check_correl_normalized(emu, 1024, fast=True)
    cases.check_autocorrel_fast(emu, 1024)
    # Correlation.rs:505-512 test_fast_normalized_correlation; :435-449 holds for the fast variant only (the plain
    # variant's n <= 32 branch has no 1/n: literal result n at lag 0)
    x = [1.0, 2.0, 3.0, 4.0]
    assert abs(nb.correl_normalized_fast(x, x, emu)[0] - 1.0) < 1e-10
    y = nb.correl_normalized_fast(x, x, emu)
    assert np.all(y >= -1.0) and np.all(y <= 1.0)
    assert abs(nb.correl_normalized(x, x, emu)[0] - 4.0) < 1e-10
    for f in (nb.correl_normalized, nb.correl_normalized_fast):
        with pytest.raises(nb.CorrelError) as ei:
            f([], [1.0], emu)
        assert ei.value.kind == nb.CorrelError.EmptyInput
        with pytest.raises(nb.CorrelError) as ei:
            f([1.0, 2.0], [1.0], emu)
        assert ei.value.kind == nb.CorrelError.LengthMismatch
        with pytest.raises(nb.CorrelError) as ei:
            f(np.ones(64), np.arange(64.0), emu)
        assert ei.value.kind == nb.CorrelError.ZeroStdDev


@pytest.mark.parametrize("n", [1, 2, 7, 32, 64, 1024, 1 << 14, 1 << 15])
def test_autocorrel_fast(emu, n):
    cases.check_autocorrel_fast(emu, n)
    if n == 2:
        # Correlation.rs:481-491 expects 10 at lag 0 (pinned) and 8 at lag 2, but the n <= 32 branch is linear
        # (Correlation.rs:294-303): lag 2 = 1*1 + 2*2 = 5 -- literal
        assert list(nb.autocorrel_fast([1.0, 2.0, 1.0, 2.0], emu))[::2] == [10.0, 5.0]
        with pytest.raises(nb.CorrelError):
            nb.autocorrel_fast([], emu)


@pytest.mark.parametrize("npoints", [1, 7, 1024, 5000])
def test_spectrum_helpers(emu, npoints):
    cases.check_spectrum(emu, npoints)
    # FFT_1.rs:305-321: power = magnitude^2
    c = O.fill_uniform(1, 0, 2 * npoints)
    assert np.max(np.abs(nb.magnitude_spectrum(c, emu) ** 2 - nb.power_spectrum(c, emu))) < 1e-10


def test_next_rows_plan_api_batched(emu):
    """Device-resident plan API for the new kinds (batched)."""
    n, cnt = 256, 3
    a = O.fill_uniform(21, 0, n * cnt) + 0.5
    b = O.fill_uniform(22, 0, n * cnt)
    out = np.zeros(4 * cnt + n * cnt)
    plan = emu.plan_create(nb.KIND_CORREL_NORM, [n], batch=cnt)
    plan.exec(a.ctypes.data, b.ctypes.data, out.ctypes.data)
    plan.destroy()
    st = out[:4 * cnt].reshape(2 * cnt, 2)
    for i in range(cnt):
        x, y = a[i * n:(i + 1) * n], b[i * n:(i + 1) * n]
        assert abs(st[i, 0] - x.mean()) < 1e-14 and abs(st[i, 1] - x.std()) < 1e-14
        assert abs(st[cnt + i, 0] - y.mean()) < 1e-14 and abs(st[cnt + i, 1] - y.std()) < 1e-14
        assert cases.rel(out[4 * cnt + i * n:4 * cnt + (i + 1) * n], O.correl_normalized(x, y)[1]) <= cases.tol(n)
    f = np.zeros(2 * cnt * (2 * n + 2))
    plan = emu.plan_create(nb.KIND_TWOFFT, [n], batch=cnt)
    plan.exec(a.ctypes.data, b.ctypes.data, f.ctypes.data)
    plan.destroy()
    per = 2 * n + 2
    for i in range(cnt):
        r1, r2 = O.twofft(a[i * n:(i + 1) * n], b[i * n:(i + 1) * n])
        assert cases.rel(f[i * per:(i + 1) * per], r1) <= cases.tol(n)
        assert cases.rel(f[(cnt + i) * per:(cnt + i + 1) * per], r2) <= cases.tol(n)


# ---------------------------------------------------------------- cosft1 / cosft2 / sinft
@pytest.mark.parametrize("n", [2, 4, 8, 64, 256, 1024, 4096, 1 << 15])
def test_cosft_sinft(emu, n):
    cases.check_cosft1(emu, n)
    cases.check_cosft2(emu, n)
    cases.check_sinft(emu, n)


def test_cosft_long_line_path_batch_and_errors(emu):
    emu.set_option("row_max_log2", 2)       # realft = c2c passes + standalone untangle
    emu.set_option("col_max_log2", 3)
    for n in (64, 2048):
        cases.check_cosft1(emu, n)
        cases.check_cosft2(emu, n)
        cases.check_sinft(emu, n)
    emu.set_option("row_max_log2", 13)
    emu.set_option("col_max_log2", 10)
    # Cos_FT2.rs:266-272: invalid isign panics
    with pytest.raises(nb.NrbError):
        nb.cosft2(np.zeros(9), 8, 0, emu)
    assert emu.cosft1(np.zeros(8), 6) == nb._lib.NRB_ERR_NOT_POW2
    # device-resident plan API, batched lines
    n, cnt = 512, 5
    y = O.fill_uniform(31, 0, cnt * (n + 2))
    refs = [O.cosft1(y[i * (n + 2):(i + 1) * (n + 2)].copy(), n) for i in range(cnt)]
    plan = emu.plan_create(nb.KIND_COSFT1, [n], batch=cnt)
    plan.exec(y.ctypes.data)
    plan.destroy()
    for i in range(cnt):
        assert cases.rel(y[i * (n + 2) + 1:(i + 1) * (n + 2)], refs[i][1:]) <= cases.tol(n)
    y = O.fill_uniform(32, 0, cnt * (n + 1))
    refs = [O.cosft2(y[i * (n + 1):(i + 1) * (n + 1)].copy(), n, 1)[1] for i in range(cnt)]
    plan = emu.plan_create(nb.KIND_COSFT2, [n], batch=cnt)
    plan.exec(y.ctypes.data, isign=1)
    plan.destroy()
    for i in range(cnt):
        assert cases.rel(y[i * (n + 1) + 1:(i + 1) * (n + 1)], refs[i][1:]) <= cases.tol(n)


# one-kernel cosft1 / cosft2 / sinft and twofft (trig_fused.cuh): every built line length, ragged last tiles, lines on
# both 16-byte alignments, against the oracle and against the multi-launch path
@pytest.mark.parametrize("n,cnt", [(16, 300), (32, 17), (64, 40), (128, 33), (256, 9), (512, 5), (1024, 5), (2048, 3), (4096, 3),
                                   (8192, 2), (16384, 2)])
def test_trig_fused_one_kernel(emu, n, cnt):
    before = _emu_count(emu, 5)
    cases.check_trig_batch(emu, n, cnt)
    assert _emu_count(emu, 5) == before + 4          # cosft1, cosft2 (+1 / -1), sinft: one launch each


@pytest.mark.parametrize("n,cnt", [(8, 300), (16, 130), (32, 70), (64, 33), (128, 17), (256, 9), (512, 5), (1024, 3), (2048, 2),
                                   (4096, 2), (8192, 2)])
def test_twofft_fused_one_kernel(emu, n, cnt):
    before = _emu_count(emu, 5)
    cases.check_twofft_plan(emu, n, cnt)
    assert _emu_count(emu, 5) == before + 1


def test_trig_fused_limits_and_unaligned_signals(emu):
    # lines the one-kernel path is not built for keep the multi-launch programs
    for kind, n in ((nb.KIND_COSFT1, 8), (nb.KIND_COSFT1, 1 << 15), (nb.KIND_TWOFFT, 4), (nb.KIND_TWOFFT, 1 << 14)):
        plan = emu.plan_create(kind, [n], batch=2)
        assert plan.num_launches(1) > 1
        plan.destroy()
    # twofft on signals that are only 8-byte aligned (plan API, caller's pointers)
    n, cnt = 64, 3
    a, b = O.fill_uniform(41, 0, n * cnt + 1)[1:], O.fill_uniform(42, 0, n * cnt + 1)[1:]
    f = np.zeros(2 * cnt * (2 * n + 2))
    plan = emu.plan_create(nb.KIND_TWOFFT, [n], batch=cnt)
    assert plan.num_launches(1) == 1
    plan.exec(a.ctypes.data, b.ctypes.data, f.ctypes.data)
    plan.destroy()
    per = 2 * n + 2
    for i in range(cnt):
        r1, r2 = O.twofft(a[i * n:(i + 1) * n], b[i * n:(i + 1) * n])
        assert cases.rel(f[i * per:(i + 1) * per], r1) <= cases.tol(n)
        assert cases.rel(f[(cnt + i) * per:(cnt + i + 1) * per], r2) <= cases.tol(n)


def test_device_resident_chain(emu):
    """SURVEY.md 8f N1 (under emulation 'device' memory is host memory)."""
    cases.check_device_resident_chain(emu)
    cases.check_device_resident_chain(emu, (4, 4, 32))
    a = O.fill_uniform(1, 0, 64)
    b = O.fill_uniform(2, 0, 64)
    pa, pb = emu.device_alloc(512), emu.device_alloc(512)
    emu.upload(pa, a)
    emu.upload(pb, b)
    emu.complex_multiply_device(pa, pb, 32, True, 0.5)
    out = np.empty(64)
    emu.download(out, pa)
    emu.stream_synchronize()
    za, zb = a[0::2] + 1j * a[1::2], b[0::2] + 1j * b[1::2]
    assert np.allclose(out[0::2] + 1j * out[1::2], za * np.conj(zb) * 0.5, rtol=1e-15, atol=0)
    emu.device_free(pa)
    emu.device_free(pb)


@pytest.mark.parametrize("n", [128, 256, 512, 1024])
def test_convlv_correl_transposed_spectrum_path(emu, n):
    """Lines longer than a tile: two passes per transform with the spectrum left in transposed order
    (AUX_SPECTRAL_ZT), against the oracle, and the same answers as the three-factor natural-order plan."""
    emu.set_option("row_max_log2", 5)
    emu.set_option("col_max_log2", 4)
    for flag in (1, 0):
        emu.set_option("conv_transposed", flag)
        cases.check_convlv(emu, n, 9)
        cases.check_correl(emu, n)
        cases.check_autocorrel_fast(emu, n)
        cases.check_correl_normalized(emu, n)
    emu.set_option("conv_transposed", 1)
    plan = emu.plan_create(nb.KIND_CONVLV, [n, 9], batch=3)
    sig, taps, ans = O.fill_uniform(1, 0, 3 * n), np.ones(9), np.zeros(3 * n)
    names = [nm for nm, _, _ in plan.profile(sig.ctypes.data, taps.ctypes.data, ans.ctypes.data)]
    plan.destroy()
    assert sum(nm.startswith("fft_") for nm in names[-5:]) == 4 and "aux_spectral_zt" in names, names


def test_convlv_default_plan_uses_two_passes_per_transform(emu):
    plan = emu.plan_create(nb.KIND_CORREL, [1 << 15], batch=1)
    assert plan.num_launches(1) == 5          # second operand 2 passes; strided pass, fused middle, strided pass (conv_fused_mid)
    plan.destroy()
    emu.set_option("conv_fused_mid", 0)
    plan = emu.plan_create(nb.KIND_CORREL, [1 << 15], batch=1)
    assert plan.num_launches(1) == 7          # 2 + 2 forward, spectral, 2 inverse
    plan.destroy()
    emu.set_option("conv_transposed", 0)
    plan = emu.plan_create(nb.KIND_CORREL, [1 << 15], batch=1)
    assert plan.num_launches(1) == 7          # 2^14 complex = 2 natural-order passes as well
    plan.destroy()


# ---- addressing fast path (fft_stage SIMPLE) and the big-tile pass (fft_pass2.cuh) ----
def _emu_count(emu, which):
    import ctypes
    emu.L.nrb_emu_launch_count.restype = ctypes.c_long
    return emu.L.nrb_emu_launch_count(which)


def _ab_equal(emu, run, option, a, b):
    """The same call under two settings of an option must give bit-identical results (same arithmetic, other addressing)."""
    emu.set_option(option, a)
    ra = run()
    emu.set_option(option, b)
    rb = run()
    assert np.array_equal(ra, rb)


@pytest.mark.parametrize("shape", [(8192,), (4, 8192), (512, 8), (1024, 32), (2, 512, 16), (1024, 8, 2)])
def test_simple_addressing_path_is_taken_and_bit_identical(emu, shape):
    n = int(np.prod(shape))
    x = cases.gen(77, 2 * n)

    def run():
        y = x.copy()
        nb.fourn(y, list(shape), len(shape), 1, emu)
        return y

    before = _emu_count(emu, 3)
    _ab_equal(emu, run, "simple_addr", 1, 0)
    assert _emu_count(emu, 3) > before          # the lengths built with the fast path (8192 contiguous, 512 / 1024 strided)
    cases.check_fourn(emu, shape)


def test_simple_addressing_with_four_step_twiddle_and_ragged_tiles(emu):
    # strided 2^18-point axis = 512 x 512: the first pass is a PLAIN strided pass with the four-step twiddle and a
    # transposed store (tw_on), 4 lines -> a ragged tile (4 of 8 lines exist)
    emu.set_option("col_max_log2", 9)
    before = _emu_count(emu, 3)
    cases.check_fourn(emu, (1 << 18, 4))
    assert _emu_count(emu, 3) > before
    x = cases.gen(78, 2 * (1 << 18) * 4)

    def run():
        y = x.copy()
        nb.fourn(y, [1 << 18, 4], 2, -1, emu)
        return y

    _ab_equal(emu, run, "simple_addr", 1, 0)


@pytest.mark.parametrize("nn,count", [(8192, 7), (4096, 9), (2048, 11)])
def test_big_tile_pass_contiguous_lines(emu, nn, count):
    emu.set_option("big_row_mask", (1 << 11) | (1 << 12) | (1 << 13))
    before = _emu_count(emu, 2)
    cases.check_four1_batch(emu, nn, count)          # more tiles than emulated persistent CTAs, ragged last tile
    cases.check_four1(emu, nn)
    assert _emu_count(emu, 2) > before


@pytest.mark.parametrize("shape", [(512, 16), (1024, 8), (2, 512, 8), (1024, 64)])
def test_big_tile_pass_strided_lines(emu, shape):
    emu.set_option("big_col_mask", (1 << 9) | (1 << 10))
    before = _emu_count(emu, 2)
    cases.check_fourn(emu, shape)
    assert _emu_count(emu, 2) > before


def test_big_tile_pass_transposing_four_step(emu):
    emu.set_option("big_col_mask", (1 << 9) | (1 << 10))
    before = _emu_count(emu, 2)
    cases.check_four1(emu, 1 << 20)                  # XPOSE 1024 + strided 1024
    cases.check_four1(emu, 1 << 19)                  # XPOSE 1024 + strided 512
    assert _emu_count(emu, 2) - before >= 16


def test_twofft_processor_batch_and_helpers(emu):
    cases.check_twofft_batch(emu, [64, 64, 256, 64, 1024, 8])
    with pytest.raises(AssertionError):                         # FFT_2.rs:6
        nb.TwoFFTProcessor(emu).process_batch([(np.zeros(8), np.zeros(8), np.zeros(17), np.zeros(18))])
    with pytest.raises(AssertionError):                         # FFT_2.rs:377
        nb.combine_real_imag(np.zeros(3), np.zeros(4))
    assert emu.twofft_batch([], [], [], []) == 0                # empty batch: nothing to do


@pytest.mark.parametrize("shp", [(8, 8, 8), (4, 16, 8), (16, 8, 32), (2, 2, 2), (32, 64, 16)])
def test_rlft3_speq_passes_on_the_side_lane(emu, shp):
    """speq_side = 1 reorders the program (z, speq passes on lane 1, y, x); same results, and the option is a no-op
    for the lane-less executions of the profiler."""
    emu.set_option("speq_side", 1)
    cases.check_rlft3(emu, shp)
    plan = emu.plan_create(nb.KIND_RLFT3, list(shp))
    assert plan.num_launches(1) == plan.num_launches(-1)


@pytest.mark.parametrize("row_max,n", [(4, 64), (4, 1024), (6, 256), (6, 4096), (12, 1 << 14), (12, 1 << 15)])
def test_conv_fused_middle_kernel(emu, row_max, n):
    """conv_mid.cuh: contiguous forward pass + spectral step + contiguous inverse pass of a row pair in one kernel.
    Rows of 16 / 64 / 4096 points (the built lengths), F = 2 (only the self-paired rows) and larger."""
    emu.set_option("row_max_log2", row_max)
    emu.set_option("conv_fused_mid", 1)
    before = _emu_count(emu, 4)
    cases.check_convlv(emu, n, min(n, 33))            # multiply (both paddings) and divide
    cases.check_correl(emu, n)
    cases.check_autocorrel_fast(emu, n)
    sigs = [cases.gen(50 + b, n) for b in range(3)]
    r = cases.gen(60, 7) / 8
    for g, a in zip(nb.convlv_batch(sigs, r, 1, 0, emu), sigs):
        assert cases.rel(g, O.convlv(a, r, 1, 0)[1]) <= cases.tol(n)
    pairs = [(cases.gen(70 + b, n), cases.gen(80 + b, n)) for b in range(3)]
    for g, (a, b) in zip(nb.correl_batch(pairs, emu), pairs):
        assert cases.rel(g, O.correl(a, b)[1]) <= cases.tol(n)
    assert _emu_count(emu, 4) - before == 7           # every long-line program went through the fused kernel


def test_conv_fused_middle_falls_back_when_not_built(emu):
    emu.set_option("row_max_log2", 5)                 # rows of 32 points: no fused kernel for that length
    emu.set_option("conv_fused_mid", 1)
    before = _emu_count(emu, 4)
    cases.check_convlv(emu, 1024, 9)
    cases.check_correl(emu, 1024)
    assert _emu_count(emu, 4) == before


def test_thread_contexts_are_released_at_thread_exit_and_by_shutdown(emu):
    """Every calling thread owns a stream and staging buffers (api.cpp ThreadCtx): short-lived threads must be able to
    come and go, nrb_shutdown releases what is left, and the library stays usable afterwards."""
    import threading
    errs = []

    def work(i):
        try:
            x = cases.gen(i, 2 * 256)
            ref = O.four1(x.copy(), 256, 1)
            nb.four1(x, 256, 1, emu)
            assert cases.rel(x, ref) <= cases.tol(256)
        except Exception as e:      # noqa: BLE001
            errs.append(e)

    for _ in range(4):
        ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    assert not errs
    emu.shutdown()
    cases.check_four1(emu, 64)


@pytest.mark.parametrize("count", [4, 9, 23])
def test_batch_calls_pipelined_in_chunks(emu, count):
    """Host-slice batch calls run in chunks over three streams (H2D | transforms | D2H): same results as the oracle per
    signal for ragged chunk counts (a shorter last chunk has its own plan), for fft_batch, realft batches, convlv_batch
    (one response for all chunks) and correl_batch (one second operand per pair), and the same as the one-shot path."""
    emu.set_option("pipeline_min_kb", 4)      # 4 KiB chunks: these small batches split into several
    before = emu.multi_device_calls(2)
    nn = 256
    arrs = [cases.gen(10 + b, 2 * nn) for b in range(count)]
    refs = [O.four1(a.copy(), nn, -1) for a in arrs]
    nb.FFTProcessor(emu).fft_batch(arrs, -1)
    for a, r in zip(arrs, refs):
        assert cases.rel(a, r) <= cases.tol(nn)
    n, m = 1024, 9
    sigs = [cases.gen(40 + b, n) for b in range(count)]
    resp = cases.gen(99, m)
    piped = nb.convlv_batch(sigs, resp, 1, 0, emu)
    for sg, o in zip(sigs, piped):
        assert cases.rel(o, O.convlv(sg, resp, 1)[1]) <= cases.tol(n)
    pairs = [(cases.gen(60 + b, n), cases.gen(80 + b, n)) for b in range(count)]
    pc = nb.correl_batch(pairs, emu)
    for (a, b), o in zip(pairs, pc):
        assert cases.rel(o, O.correl(a, b)[1]) <= cases.tol(n)
    reals = [cases.gen(120 + b, n) for b in range(count)]
    rrefs = [O.realft(r.copy(), n, 1) for r in reals]
    nb.RealFTProcessor(emu).process_batch([(r, n, 1) for r in reals])
    for r, rr in zip(reals, rrefs):
        assert cases.rel(r, rr) <= cases.tol(n)
    assert emu.multi_device_calls(2) - before == 4      # all four batch calls took the pipeline
    emu.set_option("pipeline_batches", 0)
    for a, b in zip(nb.convlv_batch(sigs, resp, 1, 0, emu), piped):
        assert np.array_equal(a, b)
    for a, b in zip(nb.correl_batch(pairs, emu), pc):
        assert np.array_equal(a, b)


def test_plans_survive_shutdown(emu):
    """nrb_shutdown drops the cache's twiddle tables; a plan created earlier holds its own references (plan.h TableRef),
    so executing it afterwards -- with other sizes planned in between, which reallocate tables -- is still exact."""
    nn = 512
    plan = emu.plan_create(nb.KIND_FOUR1, [nn])
    emu.shutdown()
    for other in (64, 256, 1024, 4096):
        cases.check_four1(emu, other)
    x = cases.gen(77, 2 * nn)
    ref = O.four1(x.copy(), nn, 1)
    d = emu.device_alloc(16 * nn)
    emu.upload(d, x)
    plan.exec(d, isign=1)
    out = np.empty_like(x)
    emu.download(out, d)
    emu.device_free(d)
    plan.destroy()
    assert cases.rel(out, ref) <= cases.tol(nn)


def test_twofft_batch_argument_errors(emu):
    import ctypes
    from numrs_b200 import _lib
    dp = ctypes.POINTER(ctypes.c_double)
    a = np.zeros(8)
    f = np.zeros(18)
    one = lambda x: (dp * 1)(x.ctypes.data_as(dp))      # noqa: E731
    null = (dp * 1)(None)
    call = emu.L.nrb_twofft_batch
    assert call(one(a), one(a), 1, 0, one(f), one(f)) == _lib.NRB_ERR_EMPTY_INPUT        # n = 0
    assert call(one(a), one(a), 1, 6, one(f), one(f)) == _lib.NRB_ERR_NOT_POW2
    assert call(one(a), one(a), 0, 8, one(f), one(f)) == 0                               # empty batch
    assert call(None, one(a), 1, 8, one(f), one(f)) == _lib.NRB_ERR_EMPTY_INPUT
    assert call(one(a), null, 1, 8, one(f), one(f)) == _lib.NRB_ERR_EMPTY_INPUT          # a null slice inside the batch
    assert "null" in emu.last_error()
