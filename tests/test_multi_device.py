"""CPU tier: ONE host-slice call spread over several devices inside one process (numrs_b200/csrc/multi.cpp).

The emulated backend pretends to have NRB_EMU_DEVICES devices (all host memory), so the scatter / slab programs /
fused exchange / gather of nrb_rlft3 and 3-D nrb_fourn, and the batch sharding of the *_batch entry points, run here
with the same host code as on a multi-GPU box; results are compared element-wise with the oracle."""
import os

import numpy as np
import pytest

import cases
import numrs_b200 as nb
import oracle as O


@pytest.fixture
def multi(emu):
    os.environ["NRB_EMU_DEVICES"] = "8"
    yield emu
    emu.set_option("num_devices", 1)
    emu.set_option("shard_min_kb", 16384)
    os.environ["NRB_EMU_DEVICES"] = "1"


@pytest.mark.parametrize("shape,G", [((8, 8, 8), 2), ((16, 16, 8), 4), ((8, 16, 32), 8), ((32, 8, 4), 0), ((16, 16, 16), 3),
                                     ((2, 2, 4), 4), ((4, 2, 8), 4), ((8, 16, 128), 4), ((8, 8, 256), 8)])   # the last two: push + pull split on
def test_rlft3_host_call_over_several_devices(multi, shape, G):
    """nrb_rlft3 on whole host arrays (Real_FT3.rs:8 call shape) with num_devices = G: forward spectrum and speq plane
    element-wise against the oracle, then the inverse and the round trip.  G = 0 means every visible device; G = 3
    rounds down to 2; shapes with fewer rows than devices fall back to one device."""
    multi.set_option("num_devices", G)
    used = multi.num_devices_in_use()
    assert used == {0: 8, 3: 2}.get(G, G)
    before = multi.multi_device_calls(0)
    n = int(np.prod(shape))
    x = O.fill_uniform(1006, 0, n).reshape(shape)
    rd, rs = O.rlft3(x.copy(), np.zeros((shape[0], 2 * shape[1])), 1)
    d, s = x.copy(), np.zeros((shape[0], 2 * shape[1]))
    nb.rlft3(d, s, *shape, 1, multi)
    assert cases.rel(d, rd) <= cases.tol(n) and cases.rel(s, rs) <= cases.tol(n)
    nb.rlft3(d, s, *shape, -1, multi)
    assert cases.rel(d * (2.0 / n), x) <= cases.tol(n)
    # the inverse of a spectrum that is not Hermitian-consistent must follow NR too (plan.cpp dc_inverse_speq)
    g = cases.gen(5, n).reshape(shape)
    gs = cases.gen(6, 2 * shape[0] * shape[1]).reshape(shape[0], 2 * shape[1])
    want, _ = O.rlft3(g.copy(), gs.copy(), -1)
    got, gs2 = g.copy(), gs.copy()
    nb.rlft3(got, gs2, *shape, -1, multi)
    assert cases.rel(got, want) <= cases.tol(n)
    # no silent single-device fallback: all three calls took the slab path whenever the shape has a row per device
    slab_ok = used <= shape[0] and used <= shape[1]
    assert multi.multi_device_calls(0) - before == (3 if slab_ok else 0)


@pytest.mark.parametrize("shape,G", [((8, 8, 8), 2), ((16, 16, 4), 4), ((8, 16, 32), 8), ((64, 128, 32), 8), ((2, 4, 8), 4), ((8, 16, 64), 4)])
def test_fourn3d_host_call_over_several_devices(multi, shape, G):
    multi.set_option("num_devices", G)
    before = multi.multi_device_calls(0)
    n = int(np.prod(shape))
    for isign in (1, -1):
        x = O.fill_uniform(1008, 0, 2 * n)
        ref = O.fourn(x.copy(), list(shape), isign)
        nb.fourn(x, list(shape), 3, isign, multi)
        assert cases.rel(x, ref) <= cases.tol(n), isign
    assert multi.multi_device_calls(0) - before == (2 if G <= shape[0] and G <= shape[1] else 0)


def test_multi_device_plans_follow_shape_changes_and_shutdown(multi):
    multi.set_option("num_devices", 4)
    for shape in ((8, 8, 8), (16, 8, 4), (8, 8, 8), (4, 16, 16)):       # more shapes than the 2-entry plan cache
        n = int(np.prod(shape))
        x = O.fill_uniform(3, 0, n).reshape(shape)
        rd, rs = O.rlft3(x.copy(), np.zeros((shape[0], 2 * shape[1])), 1)
        d, s = x.copy(), np.zeros((shape[0], 2 * shape[1]))
        nb.rlft3(d, s, *shape, 1, multi)
        assert cases.rel(d, rd) <= cases.tol(n) and cases.rel(s, rs) <= cases.tol(n)
    multi.shutdown()
    cases.check_rlft3(multi, (8, 8, 8))


@pytest.mark.parametrize("G,count", [(2, 7), (4, 9), (8, 5), (4, 2)])
def test_batches_shard_over_devices(multi, G, count):
    """fft_batch / convlv_batch / correl_batch / RealFTProcessor batches: contiguous batch ranges per device (ragged:
    count not a multiple of G, fewer signals than devices), no communication; same results as the oracle per signal."""
    multi.set_option("num_devices", G)
    multi.set_option("shard_min_kb", 0)
    before = multi.multi_device_calls(1)
    nn = 256
    arrs = [cases.gen(10 + b, 2 * nn) for b in range(count)]
    refs = [O.four1(a.copy(), nn, 1) for a in arrs]
    nb.FFTProcessor(multi).fft_batch(arrs, 1)
    for a, r in zip(arrs, refs):
        assert cases.rel(a, r) <= cases.tol(nn)
    n, m = 512, 9
    sigs = [cases.gen(40 + b, n) for b in range(count)]
    resp = cases.gen(99, m)
    outs = nb.convlv_batch(sigs, resp, 1, 0, multi)
    for sg, o in zip(sigs, outs):
        assert cases.rel(o, O.convlv(sg, resp, 1)[1]) <= cases.tol(n)
    pairs = [(cases.gen(60 + b, n), cases.gen(80 + b, n)) for b in range(count)]
    outs = nb.correl_batch(pairs, multi)
    for (a, b), o in zip(pairs, outs):
        assert cases.rel(o, O.correl(a, b)[1]) <= cases.tol(n)
    reals = [cases.gen(120 + b, n) for b in range(count)]
    rrefs = [O.realft(r.copy(), n, 1) for r in reals]
    nb.RealFTProcessor(multi).process_batch([(r, n, 1) for r in reals])
    for r, rr in zip(reals, rrefs):
        assert cases.rel(r, rr) <= cases.tol(n)
    assert multi.multi_device_calls(1) - before == 4
