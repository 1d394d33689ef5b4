"""GPU tier: the REAL multi-process exchange (one process per rank, torch.distributed.run) against the oracle.

tests/test_slab.py checks the slab programs under emulation and tests/test_gpu_slab.py with all ranks simulated on one
device by explicit copies; here every rank is its own process, the receive buffers are CUDA-IPC mappings, stage-0 kernels
store into them and the epoch-flag kernels order the stages -- the code path bench.py --gpus N times.  The forward
spectrum of every rank's slab is compared element-wise with the oracle (tests/workers/multirank_slab.py).

Shapes with nn3 >= 128 take the push + pull split of the exchange (stage 1 reads half of every block from the producer's
send buffer over the peer mapping), the others push everything.

Rank counts: 2 and 4 always (on a box with fewer GPUs the ranks share devices, which keeps every piece of the path except
the NVLink wire), 8 when the box has 8 GPUs; the NCCL exchange mode whenever there is one GPU per rank."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "workers", "multirank_slab.py")


def _ndev():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:      # noqa: BLE001
        return 0


NDEV = _ndev()
PORT = [29611]


def launch(ranks, *args, timeout=600):
    PORT[0] += 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(PORT[0]), WORKER, *args]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + "\n" + out.stderr[-4000:]
    assert "multirank_slab ok" in out.stdout, out.stdout[-2000:]
    return out.stdout


CASES = [(2, "64x128x32"), (4, "64x128x32"), (2, "256x256x256"), (4, "32x64x512")] + ([(8, "64x128x32"), (8, "256x256x256"), (8, "64x128x256")] if NDEV >= 8 else [])


@pytest.mark.parametrize("ranks,shape", CASES)
def test_rlft3_fused_exchange_forward_spectrum_vs_oracle(gpu, ranks, shape):
    launch(ranks, "--kind", "rlft3", "--mode", "fused", "--shape", shape)


@pytest.mark.parametrize("ranks,shape", [(2, "64x128x32"), (4, "64x128x32")] + ([(8, "64x128x32")] if NDEV >= 8 else []))
def test_fourn3d_fused_exchange_forward_spectrum_vs_oracle(gpu, ranks, shape):
    launch(ranks, "--kind", "fourn", "--mode", "fused", "--shape", shape)


@pytest.mark.parametrize("ranks", [2, 4])
def test_rlft3_pipelined_exchange_vs_oracle(gpu, ranks):
    launch(ranks, "--kind", "rlft3", "--mode", "fused", "--chunks", "2", "--shape", "64x128x64")


if NDEV >= 2:
    @pytest.mark.parametrize("ranks", [g for g in (2, 4, 8) if g <= NDEV])
    @pytest.mark.parametrize("kind", ["rlft3", "fourn"])
    def test_nccl_exchange_forward_spectrum_vs_oracle(gpu, ranks, kind):
        launch(ranks, "--kind", kind, "--mode", "nccl", "--shape", "64x128x32")


# ---- ONE process, several devices: the in-library multi-device path of the host-slice entry points (multi.cpp) ----
if NDEV >= 2:
    import numpy as np

    import cases
    import numrs_b200 as nb
    import oracle as O

    @pytest.fixture
    def all_devices(gpu):
        gpu.set_option("num_devices", 0)
        yield gpu
        gpu.set_option("num_devices", 1)
        gpu.set_option("shard_min_kb", 16384)

    @pytest.mark.parametrize("shape", [(64, 128, 32), (256, 256, 256), (16, 512, 1024)])
    def test_rlft3_host_call_over_all_devices_vs_oracle(all_devices, shape):
        L = all_devices
        n = int(np.prod(shape))
        before = L.multi_device_calls(0)
        x = O.fill_uniform(1006, 0, n).reshape(shape)
        rd, rs = O.rlft3(x.copy(), np.zeros((shape[0], 2 * shape[1])), 1, mt=True)
        d, s = x.copy(), np.zeros((shape[0], 2 * shape[1]))
        nb.rlft3(d, s, *shape, 1)
        assert cases.rel(d, rd) <= cases.tol(n) and cases.rel(s, rs) <= cases.tol(n)
        nb.rlft3(d, s, *shape, -1)
        assert cases.rel(d * (2.0 / n), x) <= cases.tol(n)
        assert L.multi_device_calls(0) - before == 2 and L.num_devices_in_use() >= 2

    @pytest.mark.parametrize("shape", [(64, 128, 32), (128, 128, 128)])
    def test_fourn3d_host_call_over_all_devices_vs_oracle(all_devices, shape):
        n = int(np.prod(shape))
        for isign in (1, -1):
            x = O.fill_uniform(1008, 0, 2 * n)
            ref = O.fourn(x.copy(), list(shape), isign, mt=True)
            nb.fourn(x, list(shape), 3, isign)
            assert cases.rel(x, ref) <= cases.tol(n), isign

    def test_batches_shard_over_all_devices_vs_oracle(all_devices):
        L = all_devices
        before = L.multi_device_calls(1)
        nn, cnt = 4096, 1031                       # 64 MiB, ragged over the devices
        x = O.fill_uniform(1002, 0, 2 * nn * cnt)
        y = x.copy()
        arrs = [y[2 * nn * b:2 * nn * (b + 1)] for b in range(cnt)]
        nb.FFTProcessor().fft_batch(arrs, 1)
        for b in (0, 1, 515, 516, 1030):
            assert cases.rel(arrs[b], O.four1(x[2 * nn * b:2 * nn * (b + 1)].copy(), nn, 1)) <= cases.tol(nn)
        nb.FFTProcessor().fft_batch(arrs, -1)
        assert cases.rel(y / nn, x) <= cases.tol(nn)
        n, m, c2 = 1 << 20, 4096, 5
        sigs = [O.fill_uniform(1004, i * n, n) for i in range(c2)]
        r = O.fill_uniform(1005, 0, m) / 64
        for sg, o in zip(sigs, nb.convlv_batch(sigs, r, 1)):
            assert cases.rel(o, O.convlv(sg, r, 1)[1]) <= cases.tol(n)
        tm = [np.concatenate([O.fill_uniform(1005, 0, m), np.zeros(n - m)]) for _ in range(c2)]
        for sg, t, o in zip(sigs, tm, nb.correl_batch(list(zip(sigs, tm)))):
            assert cases.rel(o, O.correl(sg, t)[1]) <= cases.tol(n)
        assert L.multi_device_calls(1) - before == 4
