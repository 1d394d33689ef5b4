#!/usr/bin/env python
"""Runs one small instance of every kernel family through the C ABI (for compute-sanitizer memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import cases  # noqa: E402
import numrs_b200 as nb  # noqa: E402

L = nb.lib()
for nn in (2, 8, 64, 256, 512, 4096, 8192, 1 << 14):
    cases.check_four1(L, nn)
cases.check_four1_batch(L, 256, 37)
for shape in ((4, 8, 2), (64, 64), (2, 1024), (2048, 4), (8, 8, 8, 4)):
    cases.check_fourn(L, shape)
for n in (2, 8, 64, 512, 4096, 1 << 15):
    cases.check_realft(L, n)
for shp in ((2, 2, 2), (8, 8, 8), (16, 8, 32), (4, 256, 256), (1, 1, 64)):
    cases.check_rlft3(L, shp)
L.set_option("fuse_zy", 1)
cases.check_rlft3(L, (4, 256, 256))
L.set_option("fuse_zy", 0)
L.set_option("row_max_log2", 5)
L.set_option("col_max_log2", 4)
cases.check_four1(L, 1 << 12)
cases.check_realft(L, 1 << 10)
cases.check_convlv(L, 1 << 10, 33)
cases.check_correl(L, 1 << 10)
L.set_option("row_max_log2", 13)
L.set_option("col_max_log2", 10)
for n, m in ((4, 2), (64, 5), (1024, 33)):
    cases.check_convlv(L, n, m)
for n in (3, 32, 64, 1024):
    cases.check_correl(L, n)
# next rows and the transposed-spectrum convolution path
for n in (1, 16, 4096, 1 << 14):
    cases.check_twofft(L, n)
for n in (5, 32, 64, 4096):
    cases.check_correl_normalized(L, n)
    cases.check_correl_normalized(L, n, fast=True)
    cases.check_autocorrel_fast(L, n)
for n in (2, 8, 256, 4096, 1 << 15):
    cases.check_cosft1(L, n)
    cases.check_cosft2(L, n)
    cases.check_sinft(L, n)
cases.check_spectrum(L, 5000)
cases.check_device_resident_chain(L)
for flag in (1, 0):
    L.set_option("conv_transposed", flag)
    cases.check_convlv(L, 1 << 15, 100)
    cases.check_correl(L, 1 << 15)
    cases.check_autocorrel_fast(L, 1 << 15)
L.set_option("conv_transposed", 1)
L.set_option("prefetch_dist", 3)
cases.check_rlft3(L, (16, 8, 32))
cases.check_four1(L, 4096)
L.set_option("prefetch_dist", -1)
# kernels added after the first sanitizer run: cheap addressing path, big-tile pass, fused convolution middle, side lane,
# batched twofft
for flag in (0, 1):
    L.set_option("simple_addr", flag)
    for shape in ((4, 8192), (512, 32), (1024, 8, 2)):
        cases.check_fourn(L, shape)
cases.check_four1_batch(L, 8192, 151)
cases.check_four1_batch(L, 2048, 301)
cases.check_fourn(L, (1024, 64))
cases.check_four1(L, 1 << 19)
L.set_option("conv_fused_mid", 1)
for n in (1 << 15, 1 << 16):
    cases.check_convlv(L, n, 100)
    cases.check_correl(L, n)
    cases.check_autocorrel_fast(L, n)
L.set_option("conv_fused_mid", 0)
L.set_option("speq_side", 1)
cases.check_rlft3(L, (16, 8, 32))
cases.check_rlft3(L, (64, 64, 64))
L.set_option("speq_side", 0)
cases.check_twofft_batch(L, [64, 4096, 64, 2])
L.set_option("speq_side", 1)
L.set_option("conv_fused_mid", 1)
# ---- round 2: TMA-fed strided passes (cp.async.bulk.tensor + mbarrier), both the one-CTA-per-tile and the persistent form
for persist in (0, 1):
    L.set_option("tma_persist", persist)
    for mask in (0x780, 0):
        L.set_option("tma_col_mask", mask)
        for shape in ((512, 64), (1024, 16), (4, 256, 32), (8, 128, 64), (64, 512, 8)):
            cases.check_fourn(L, shape)
        cases.check_rlft3(L, (16, 512, 32))
        cases.check_rlft3(L, (512, 16, 64))
L.set_option("tma_persist", 0)
L.set_option("tma_col_mask", (1 << 9) | (1 << 10))
# fused conv middle with 1024-thread CTAs (default) and on 2048-point rows
for rest in (12, 11):
    L.set_option("conv_rest_log2", rest)
    for n in (1 << 15, 1 << 16):
        cases.check_convlv(L, n, 100)
        cases.check_correl(L, n)
        cases.check_autocorrel_fast(L, n)
L.set_option("conv_rest_log2", 12)
# batch calls pipelined in chunks over three streams
L.set_option("pipeline_min_kb", 4)
arrs = [cases.gen(10 + b, 2 * 256) for b in range(9)]
nb.FFTProcessor(L).fft_batch(arrs, 1)
sig = [cases.gen(40 + b, 1024) for b in range(9)]
nb.convlv_batch(sig, cases.gen(99, 9), 1, 0, L)
nb.correl_batch([(a, cases.gen(80 + i, 1024)) for i, a in enumerate(sig)], L)
L.set_option("pipeline_min_kb", 16384)
# slab stages of rlft3 and fourn with the exchange tables (pushed, push + pull, z in two y-chunks): all ranks on this one device
import ctypes  # noqa: E402


def slab_roundtrip(kind, shape, G, pull, zc):
    L.set_option("z_chunks", zc)
    nn1, nn2, nn3 = shape
    X, Y = nn1 // G, nn2 // G
    plans = [L.slab_create(nn1, nn2, nn3, G, r, kind=kind) for r in range(G)]
    ld, sd, rb = plans[0].local_doubles(), plans[0].speq_doubles(), plans[0].recv_bytes()
    recv = [L.device_alloc(rb) for _ in range(G)]
    send = [L.device_alloc(rb) for _ in range(G)]
    slab = [L.device_alloc(8 * ld) for _ in range(G)]
    speq = [L.device_alloc(8 * max(sd, 2)) for _ in range(G)]
    x = [cases.gen(300 + r, ld) for r in range(G)]
    for r in range(G):
        L.upload(slab[r], x[r])
        plans[r].set_peers(recv)
        plans[r].set_send_peers(send if pull else None)
    scale = (2.0 if kind == "rlft3" else 1.0) / (nn1 * nn2 * nn3)
    for isign in (1, -1):
        for r in range(G):
            plans[r].stage(0, isign, slab[r], speq[r], 0, 0)
        L.stream_synchronize()
        for r in range(G):
            plans[r].stage(1, isign, slab[r], speq[r], 0, 0)
        L.stream_synchronize()
    for r in range(G):
        out = np.empty(ld)
        L.download(out, slab[r])
        L.stream_synchronize()
        assert cases.rel(out * scale, x[r]) <= cases.tol(nn1 * nn2 * nn3), (kind, shape, G, pull, zc, r)
    for p in plans:
        p.destroy()
    for b in recv + send + slab + speq:
        L.device_free(b)


for kind, shape in (("rlft3", (8, 64, 512)), ("fourn", (8, 64, 256)), ("rlft3", (16, 16, 32))):
    for pull in (False, True):
        for zc in (1, 2):
            slab_roundtrip(kind, shape, 2, pull, zc)
L.set_option("z_chunks", 1)
print("sanitize_small: all parity checks passed")
