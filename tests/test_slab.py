"""CPU tier: slab-decomposed rlft3 (multi-GPU path).  The per-rank stages run under the
TEST-ONLY kernel emulation; the exchange is (a) simulated in one process for G = 2, 4, 8 and
(b) a real torch.distributed all_to_all_single over gloo with world_size 2."""
import os
import sys

import numpy as np
import pytest

import cases
import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def slab_of(x, r, G):
    Y = x.shape[1] // G
    return np.ascontiguousarray(x[:, r * Y:(r + 1) * Y, :])


def run_simulated(L, x, G, isign, speq_in=None, fused=False, pull=False):
    """All G ranks in one process; returns per-rank (slab, speq) after one direction.
    fused=True: stage 0 stores straight into the peers' receive buffers (no explicit exchange); pull=True: the
    push + pull split (the high-z part of every block stays in the producer's send buffer and stage 1 reads it there)."""
    nn1, nn2, nn3 = x.shape
    X, Y = nn1 // G, nn2 // G
    plans = [L.slab_create(nn1, nn2, nn3, G, r) for r in range(G)]
    xd = plans[0].xchg_doubles()
    blk = xd // G
    if isign == 1:
        slabs = [slab_of(x, r, G).ravel().copy() for r in range(G)]
        speqs = [np.zeros(plans[r].speq_doubles()) for r in range(G)]
    else:
        slabs = [np.ascontiguousarray(x[r * X:(r + 1) * X]).ravel().copy() for r in range(G)]
        speqs = [np.ascontiguousarray(speq_in[r * X:(r + 1) * X]).ravel().copy() for r in range(G)]
    sends = [np.zeros(xd) for _ in range(G)]
    recvs = [np.zeros(plans[0].recv_bytes() // 8) for _ in range(G)]
    if fused:
        for r in range(G):
            plans[r].set_peers([rv.ctypes.data for rv in recvs])
            if pull:
                plans[r].set_send_peers([sv.ctypes.data for sv in sends])
    for r in range(G):
        plans[r].stage(0, isign, slabs[r].ctypes.data, speqs[r].ctypes.data, sends[r].ctypes.data, 0)
    for r in range(G):          # all-to-all: block p of rank r's send -> block r of rank p's recv
        for p in range(G):
            if not fused:
                recvs[p][r * blk:(r + 1) * blk] = sends[r][p * blk:(p + 1) * blk]
    if fused:       # flag barrier: every rank publishes epoch 1, then every rank sees all G flags
        for r in range(G):
            plans[r].barrier(0, 1)
        for r in range(G):
            plans[r].barrier(1, 1)
            assert list(recvs[r][xd:xd + G].view(np.uint64)) == [1] * G
    for r in range(G):
        plans[r].stage(1, isign, slabs[r].ctypes.data, speqs[r].ctypes.data, 0, recvs[r].ctypes.data)
    for p in plans:
        p.destroy()
    return slabs, speqs


@pytest.mark.parametrize("eighths", [4, 1, 8])
@pytest.mark.parametrize("shape,G", [((8, 8, 128), 2), ((16, 16, 256), 4), ((8, 16, 128), 8), ((8, 8, 32), 2)])
def test_slab_push_pull_exchange(emu, shape, G, eighths):
    """The push + pull split of the fused exchange: `eighths` / 8 of the z range of every block is left in the
    producer's send buffer and read from there by the consumer's stage 1; same spectrum as the oracle, element-wise,
    and the round trip.  (N3 < 64: the split switches itself off and everything is pushed.)"""
    emu.set_option("pull_eighths", eighths)
    nn1, nn2, nn3 = shape
    X = nn1 // G
    x = O.fill_uniform(1006, 0, nn1 * nn2 * nn3).reshape(shape)
    rd, rs = O.rlft3(x.copy(), np.zeros((nn1, 2 * nn2)), 1)
    slabs, speqs = run_simulated(emu, x, G, 1, fused=True, pull=True)
    for r in range(G):
        assert cases.rel(slabs[r], rd[r * X:(r + 1) * X]) <= cases.tol(x.size), (r, "data")
        assert cases.rel(speqs[r], rs[r * X:(r + 1) * X]) <= cases.tol(x.size), (r, "speq")
    back, _ = run_simulated(emu, rd, G, -1, rs, fused=True, pull=True)
    for r in range(G):
        assert cases.rel(back[r] * (2.0 / x.size), slab_of(x, r, G)) <= cases.tol(x.size), (r, "round trip")
    emu.set_option("pull_eighths", 4)


@pytest.mark.parametrize("shape,G", [((4, 64, 512), 2), ((8, 32, 1024), 2), ((8, 64, 256), 4)])
def test_slab_z_pass_in_two_y_chunks(emu, shape, G):
    """z_chunks = 2: the z pass and the exchange pass next to it run as two halves of the local y rows (the z pass of one
    half on the side lane beside the exchange pass of the other); every launch covers a tile subset
    (PassParams::tile_run).  rlft3 and fourn, pushed and push + pull, against the oracle element-wise."""
    emu.set_option("z_chunks", 2)
    nn1, nn2, nn3 = shape
    X, Y = nn1 // G, nn2 // G
    plan = emu.slab_create(nn1, nn2, nn3, G, 0)
    assert plan.num_launches(1) == 8      # z0, z1, speq x, x0, x1 | barrier | y, speq y
    plan.destroy()
    x = O.fill_uniform(1006, 0, nn1 * nn2 * nn3).reshape(shape)
    rd, rs = O.rlft3(x.copy(), np.zeros((nn1, 2 * nn2)), 1)
    for pull in (False, True):
        slabs, speqs = run_simulated(emu, x, G, 1, fused=True, pull=pull)
        for r in range(G):
            assert cases.rel(slabs[r], rd[r * X:(r + 1) * X]) <= cases.tol(x.size), (r, "data")
            assert cases.rel(speqs[r], rs[r * X:(r + 1) * X]) <= cases.tol(x.size), (r, "speq")
        back, _ = run_simulated(emu, rd, G, -1, rs, fused=True, pull=pull)
        for r in range(G):
            assert cases.rel(back[r] * (2.0 / x.size), slab_of(x, r, G)) <= cases.tol(x.size), (r, "round trip")
    shp = (nn1, nn2, nn3 // 2)
    n = int(np.prod(shp))
    xf = O.fill_uniform(1008, 0, 2 * n)
    ref = O.fourn(xf.copy(), list(shp), 1).view(np.complex128).reshape(shp)
    z = xf.view(np.complex128).reshape(shp)
    out = run_simulated_fourn(emu, z, G, 1, 1)
    for r in range(G):
        assert cases.rel(out[r].view(np.float64), np.ascontiguousarray(ref[r * X:(r + 1) * X]).view(np.float64)) <= cases.tol(n)
    back = run_simulated_fourn(emu, ref, G, -1, 1)
    for r in range(G):
        assert cases.rel(back[r].view(np.float64) / n, np.ascontiguousarray(z[:, r * Y:(r + 1) * Y, :]).view(np.float64)) <= cases.tol(n)


@pytest.mark.parametrize("side", [0, 1])
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("shape,G", [((8, 8, 8), 2), ((16, 16, 8), 4), ((8, 16, 32), 8), ((32, 8, 4), 8), ((4, 4, 2), 4),
                                     ((16, 16, 16), 1)])
def test_slab_simulated_ranks(emu, shape, G, fused, side):
    emu.set_option("speq_side", side)      # 1: the speq-plane pass of every stage first, on the side lane
    nn1, nn2, nn3 = shape
    X, Y = nn1 // G, nn2 // G
    x = O.fill_uniform(1006, 0, nn1 * nn2 * nn3).reshape(shape)
    rd, rs = O.rlft3(x.copy(), np.zeros((nn1, 2 * nn2)), 1)
    slabs, speqs = run_simulated(emu, x, G, 1, fused=fused)
    for r in range(G):      # forward output: nn1-slabs = contiguous row ranges of the reference layout
        assert cases.rel(slabs[r], rd[r * X:(r + 1) * X]) <= cases.tol(x.size), (r, "data")
        assert cases.rel(speqs[r], rs[r * X:(r + 1) * X]) <= cases.tol(x.size), (r, "speq")
    # inverse from the spectrum: output nn2-slabs, round trip = N/2 * x
    back, _ = run_simulated(emu, rd, G, -1, rs, fused=fused)
    for r in range(G):
        assert cases.rel(back[r] * (2.0 / x.size), slab_of(x, r, G)) <= cases.tol(x.size), (r, "round trip")


def run_simulated_fourn(L, z, G, isign, chunks=1):
    """3-D complex fourn slabs, all G ranks in one process (fused exchange; chunks > 1: the pipelined pieces).
    z: complex [nn1][nn2][nn3]; isign=+1 takes nn2-slabs and returns nn1-slabs, isign=-1 the mirror image."""
    nn1, nn2, nn3 = z.shape
    X, Y = nn1 // G, nn2 // G
    plans = [L.slab_create(nn1, nn2, nn3, G, r, kind="fourn") for r in range(G)]
    assert plans[0].speq_doubles() == 0 and plans[0].local_doubles() == 2 * z.size // G
    xd = plans[0].xchg_doubles()
    if isign == 1:
        slabs = [np.ascontiguousarray(z[:, r * Y:(r + 1) * Y, :]).view(np.float64).ravel().copy() for r in range(G)]
    else:
        slabs = [np.ascontiguousarray(z[r * X:(r + 1) * X]).view(np.float64).ravel().copy() for r in range(G)]
    recvs = [np.zeros(plans[0].recv_bytes() // 8) for _ in range(G)]
    for r in range(G):
        plans[r].set_peers([rv.ctypes.data for rv in recvs])
    if chunks == 1:
        for r in range(G):
            plans[r].stage(0, isign, slabs[r].ctypes.data, 0, 0, 0)
        for r in range(G):
            plans[r].barrier(0, 1)
        for r in range(G):
            plans[r].barrier(1, 1)
            assert list(recvs[r][xd:xd + G].view(np.uint64)) == [1] * G
            plans[r].stage(1, isign, slabs[r].ctypes.data, 0, 0, 0)
    else:
        for r in range(G):
            plans[r].set_chunks(chunks)
            plans[r].stage_part(0, -1, isign, slabs[r].ctypes.data, 0)
        for c in range(chunks):
            for r in range(G):
                plans[r].stage_part(0, c, isign, slabs[r].ctypes.data, 0)
                plans[r].barrier_chunk(0, c, 1)
            for r in range(G):
                plans[r].barrier_chunk(1, c, 1)
                plans[r].stage_part(1, c, isign, slabs[r].ctypes.data, 0)
        for r in range(G):
            plans[r].stage_part(0, chunks, isign, slabs[r].ctypes.data, 0)
    for p in plans:
        p.destroy()
    return [s.view(np.complex128) for s in slabs]


@pytest.mark.parametrize("shape,G,chunks", [((8, 8, 8), 2, 1), ((16, 16, 4), 4, 1), ((8, 16, 32), 8, 1), ((64, 128, 32), 8, 1),
                                            ((4, 4, 2), 4, 1), ((8, 8, 4), 1, 1), ((8, 8, 32), 2, 2), ((16, 16, 64), 4, 4)])
def test_fourn3d_slabs_simulated_ranks(emu, shape, G, chunks):
    """Slab-decomposed 3-D complex fourn (call shape Real_FT3.rs:35) against the oracle's in-memory fourn: forward
    spectrum element-wise (nn1-slabs = row ranges of the reference layout), the inverse, and the round trip."""
    nn1, nn2, nn3 = shape
    X, Y = nn1 // G, nn2 // G
    n = nn1 * nn2 * nn3
    x = O.fill_uniform(1008, 0, 2 * n)
    ref = O.fourn(x.copy(), list(shape), 1).view(np.complex128).reshape(shape)
    z = x.view(np.complex128).reshape(shape)
    out = run_simulated_fourn(emu, z, G, 1, chunks)
    for r in range(G):
        assert cases.rel(out[r].view(np.float64), np.ascontiguousarray(ref[r * X:(r + 1) * X]).view(np.float64)) <= cases.tol(n), (r, "forward")
    refm = O.fourn(x.copy(), list(shape), -1).view(np.complex128).reshape(shape)
    outm = run_simulated_fourn(emu, z[:, :, :].copy(), G, -1, chunks)       # isign = -1: input nn1-slabs, output nn2-slabs
    for r in range(G):
        assert cases.rel(outm[r].view(np.float64), np.ascontiguousarray(refm[:, r * Y:(r + 1) * Y, :]).view(np.float64)) <= cases.tol(n), (r, "inverse")
    back = run_simulated_fourn(emu, ref, G, -1, chunks)
    for r in range(G):
        assert cases.rel(back[r].view(np.float64) / n, np.ascontiguousarray(z[:, r * Y:(r + 1) * Y, :]).view(np.float64)) <= cases.tol(n), (r, "round trip")


def run_simulated_pipelined(L, x, G, isign, chunks, speq_in=None):
    """Pipelined (z-chunked) fused exchange with all G ranks in one process.  The emulated flag wait does not
    block, so the test orders the pieces itself: stage 0 of chunk c on every rank, then stage 1 of chunk c."""
    nn1, nn2, nn3 = x.shape
    X = nn1 // G
    plans = [L.slab_create(nn1, nn2, nn3, G, r) for r in range(G)]
    xd = plans[0].xchg_doubles()
    if isign == 1:
        slabs = [slab_of(x, r, G).ravel().copy() for r in range(G)]
        speqs = [np.zeros(plans[r].speq_doubles()) for r in range(G)]
    else:
        slabs = [np.ascontiguousarray(x[r * X:(r + 1) * X]).ravel().copy() for r in range(G)]
        speqs = [np.ascontiguousarray(speq_in[r * X:(r + 1) * X]).ravel().copy() for r in range(G)]
    recvs = [np.zeros(plans[0].recv_bytes() // 8) for _ in range(G)]
    for r in range(G):
        plans[r].set_peers([rv.ctypes.data for rv in recvs])
        plans[r].set_chunks(chunks)
        plans[r].stage_part(0, -1, isign, slabs[r].ctypes.data, speqs[r].ctypes.data)
    for c in range(chunks):
        for r in range(G):
            plans[r].stage_part(0, c, isign, slabs[r].ctypes.data, speqs[r].ctypes.data)
            plans[r].barrier_chunk(0, c, 1)
        for r in range(G):
            plans[r].barrier_chunk(1, c, 1)
            assert list(recvs[r][xd + c * G:xd + (c + 1) * G].view(np.uint64)) == [1] * G
            plans[r].stage_part(1, c, isign, slabs[r].ctypes.data, speqs[r].ctypes.data)
    for r in range(G):
        plans[r].stage_part(0, chunks, isign, slabs[r].ctypes.data, speqs[r].ctypes.data)
    for p in plans:
        p.destroy()
    return slabs, speqs


@pytest.mark.parametrize("shape,G,chunks", [((8, 8, 32), 2, 2), ((16, 16, 64), 4, 4), ((8, 16, 64), 8, 2), ((16, 8, 16), 2, 1)])
def test_slab_pipelined_exchange(emu, shape, G, chunks):
    nn1, nn2, nn3 = shape
    X = nn1 // G
    x = O.fill_uniform(1006, 0, nn1 * nn2 * nn3).reshape(shape)
    rd, rs = O.rlft3(x.copy(), np.zeros((nn1, 2 * nn2)), 1)
    if chunks == 1:     # chunks = 1 switches the pipeline off again
        plan = emu.slab_create(nn1, nn2, nn3, G, 0)
        plan.set_chunks(1)
        with pytest.raises(Exception):
            plan.stage_part(0, 0, 1, 0, 0)
        plan.destroy()
        return
    slabs, speqs = run_simulated_pipelined(emu, x, G, 1, chunks)
    for r in range(G):
        assert cases.rel(slabs[r], rd[r * X:(r + 1) * X]) <= cases.tol(x.size), (r, "data")
        assert cases.rel(speqs[r], rs[r * X:(r + 1) * X]) <= cases.tol(x.size), (r, "speq")
    back, _ = run_simulated_pipelined(emu, rd, G, -1, chunks, rs)
    for r in range(G):
        assert cases.rel(back[r] * (2.0 / x.size), slab_of(x, r, G)) <= cases.tol(x.size), (r, "round trip")
    # more chunks than complex z values, or not a power of two
    plan = emu.slab_create(nn1, nn2, nn3, G, 0)
    import numrs_b200 as nb
    for bad in (nn3, 3):
        with pytest.raises(nb.NrbError):
            plan.set_chunks(bad)
    plan.destroy()


def run_simulated_dma(L, x, G, isign, chunks, speq_in=None):
    """DMA exchange, step-wise, all G ranks in one process: stage 0 of chunk c writes each rank's send buffer
    (chunk-major blocks), the test plays the copy engines (piece (peer, chunk) -> block `rank` of the peer's
    receive buffer), then stage 1 of chunk c."""
    nn1, nn2, nn3 = x.shape
    X, Y, N3 = nn1 // G, nn2 // G, nn3 // 2
    plans = [L.slab_create(nn1, nn2, nn3, G, r) for r in range(G)]
    xd = plans[0].xchg_doubles()
    blk, spq, piece = 2 * X * Y * (N3 + 1), 2 * X * Y * N3, 2 * X * Y * (N3 // chunks)      # in doubles
    if isign == 1:
        slabs = [slab_of(x, r, G).ravel().copy() for r in range(G)]
        speqs = [np.zeros(plans[r].speq_doubles()) for r in range(G)]
    else:
        slabs = [np.ascontiguousarray(x[r * X:(r + 1) * X]).ravel().copy() for r in range(G)]
        speqs = [np.ascontiguousarray(speq_in[r * X:(r + 1) * X]).ravel().copy() for r in range(G)]
    sends = [np.full(xd, np.nan) for _ in range(G)]
    recvs = [np.full(xd, np.nan) for _ in range(G)]
    for r in range(G):
        plans[r].set_dma(chunks)
        plans[r].stage_part_xchg(0, -1, isign, slabs[r].ctypes.data, speqs[r].ctypes.data, 0)
    for c in range(chunks):
        for r in range(G):
            plans[r].stage_part_xchg(0, c, isign, slabs[r].ctypes.data, speqs[r].ctypes.data, sends[r].ctypes.data)
        for r in range(G):
            for p in range(G):
                recvs[p][r * blk + c * piece:r * blk + (c + 1) * piece] = sends[r][p * blk + c * piece:p * blk + (c + 1) * piece]
                if c == 0:
                    recvs[p][r * blk + spq:(r + 1) * blk] = sends[r][p * blk + spq:(p + 1) * blk]
        for r in range(G):
            plans[r].stage_part_xchg(1, c, isign, slabs[r].ctypes.data, speqs[r].ctypes.data, recvs[r].ctypes.data)
    for r in range(G):
        plans[r].stage_part_xchg(0, chunks, isign, slabs[r].ctypes.data, speqs[r].ctypes.data, 0)
    for p in plans:
        p.destroy()
    return slabs, speqs


@pytest.mark.parametrize("shape,G,chunks", [((8, 8, 32), 2, 2), ((16, 16, 64), 4, 4), ((8, 16, 64), 8, 1), ((16, 8, 16), 2, 8)])
def test_slab_dma_exchange(emu, shape, G, chunks):
    nn1, nn2, nn3 = shape
    X = nn1 // G
    x = O.fill_uniform(1006, 0, nn1 * nn2 * nn3).reshape(shape)
    rd, rs = O.rlft3(x.copy(), np.zeros((nn1, 2 * nn2)), 1)
    slabs, speqs = run_simulated_dma(emu, x, G, 1, chunks)
    for r in range(G):
        assert cases.rel(slabs[r], rd[r * X:(r + 1) * X]) <= cases.tol(x.size), (r, "data")
        assert cases.rel(speqs[r], rs[r * X:(r + 1) * X]) <= cases.tol(x.size), (r, "speq")
    back, _ = run_simulated_dma(emu, rd, G, -1, chunks, rs)
    for r in range(G):
        assert cases.rel(back[r] * (2.0 / x.size), slab_of(x, r, G)) <= cases.tol(x.size), (r, "round trip")


def test_slab_dma_one_call_single_rank(emu):
    """nrb_slab_exec_dma (streams, events, copies, per-chunk flags in one call) with world size 1, where the
    emulated in-order execution is a valid schedule."""
    shape = (8, 16, 32)
    nn1, nn2, nn3 = shape
    x = O.fill_uniform(1006, 0, nn1 * nn2 * nn3).reshape(shape)
    rd, rs = O.rlft3(x.copy(), np.zeros((nn1, 2 * nn2)), 1)
    plan = emu.slab_create(nn1, nn2, nn3, 1, 0)
    recv = np.zeros(plan.recv_bytes() // 8)
    plan.set_peers([recv.ctypes.data])
    plan.set_dma(4)
    slab, speq = x.ravel().copy(), np.zeros(plan.speq_doubles())
    plan.exec_dma(1, slab.ctypes.data, speq.ctypes.data, 1)
    assert cases.rel(slab, rd) <= cases.tol(x.size) and cases.rel(speq, rs) <= cases.tol(x.size)
    plan.exec_dma(-1, slab.ctypes.data, speq.ctypes.data, 2)
    assert cases.rel(slab * (2.0 / x.size), x) <= cases.tol(x.size)
    plan.destroy()


def test_slab_rejects_bad_rank_counts(emu):
    import numrs_b200 as nb
    for args in ((8, 8, 8, 3, 0), (8, 8, 8, 16, 0), (8, 8, 8, 2, 2), (8, 6, 8, 2, 0)):
        with pytest.raises(nb.NrbError):
            emu.slab_create(*args)


def _gloo_worker(rank, world, port, shape, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from numrs_b200 import _lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = _lib.Library(os.path.join(ROOT, "tests", "emu", "libnrb_emu.so"))
    nn1, nn2, nn3 = shape
    x = O.fill_uniform(1006, 0, nn1 * nn2 * nn3).reshape(shape)
    plan = L.slab_create(nn1, nn2, nn3, world, rank)
    slab = torch.from_numpy(slab_of(x, rank, world).ravel().copy())
    speq = torch.zeros(plan.speq_doubles(), dtype=torch.float64)
    send = torch.zeros(plan.xchg_doubles(), dtype=torch.float64)
    recv = torch.zeros(plan.xchg_doubles(), dtype=torch.float64)
    for isign in (1, -1):
        plan.stage(0, isign, slab.data_ptr(), speq.data_ptr(), send.data_ptr(), 0)
        dist.all_to_all_single(recv, send)
        plan.stage(1, isign, slab.data_ptr(), speq.data_ptr(), 0, recv.data_ptr())
        if isign == 1:
            fwd, fsp = slab.numpy().copy(), speq.numpy().copy()
    q.put((rank, fwd, fsp, slab.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_gloo_world_size_2(emu):
    import torch.multiprocessing as mp
    shape, world = (8, 16, 16), 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, shape, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, fwd, fsp, back = q.get(timeout=120)
        res[rank] = (fwd, fsp, back)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nn1, nn2, nn3 = shape
    x = O.fill_uniform(1006, 0, nn1 * nn2 * nn3).reshape(shape)
    rd, rs = O.rlft3(x.copy(), np.zeros((nn1, 2 * nn2)), 1)
    X = nn1 // world
    for r in range(world):
        fwd, fsp, back = res[r]
        assert cases.rel(fwd, rd[r * X:(r + 1) * X]) <= cases.tol(x.size)
        assert cases.rel(fsp, rs[r * X:(r + 1) * X]) <= cases.tol(x.size)
        assert cases.rel(back * (2.0 / x.size), slab_of(x, r, world)) <= cases.tol(x.size)
