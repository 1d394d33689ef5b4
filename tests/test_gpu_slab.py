"""GPU tier: slab-decomposed rlft3 stages (two-level exchange addressing) on the CUDA library.
All G ranks are simulated on ONE device (stage kernels on the GPU, exchange by device copies); the
multi-process NCCL path is exercised by bench.py --gpus N and tests/test_slab.py (gloo)."""
import numpy as np
import pytest

import cases
import oracle as O

pytestmark = pytest.mark.gpu


def run_direction(L, torch, x_slabs, speqs, shape, G, isign):
    nn1, nn2, nn3 = shape
    plans = [L.slab_create(nn1, nn2, nn3, G, r) for r in range(G)]
    xd = plans[0].xchg_doubles()
    blk = xd // G
    st = torch.cuda.current_stream().cuda_stream
    sends = [torch.zeros(xd, dtype=torch.float64, device="cuda") for _ in range(G)]
    recvs = [torch.zeros(xd, dtype=torch.float64, device="cuda") for _ in range(G)]
    for r in range(G):
        plans[r].stage(0, isign, x_slabs[r].data_ptr(), speqs[r].data_ptr(), sends[r].data_ptr(), 0, st)
    for r in range(G):
        for p in range(G):
            recvs[p][r * blk:(r + 1) * blk] = sends[r][p * blk:(p + 1) * blk]
    for r in range(G):
        plans[r].stage(1, isign, x_slabs[r].data_ptr(), speqs[r].data_ptr(), 0, recvs[r].data_ptr(), st)
    torch.cuda.synchronize()
    for p in plans:
        p.destroy()


@pytest.mark.parametrize("shape,G", [((8, 8, 8), 2), ((64, 64, 64), 2), ((64, 128, 32), 4), ((128, 128, 128), 8),
                                     ((256, 256, 256), 8), ((512, 512, 64), 8)])
def test_slab_stages_simulated_ranks(gpu, shape, G):
    import torch
    nn1, nn2, nn3 = shape
    X, Y = nn1 // G, nn2 // G
    x = O.fill_uniform(1006, 0, nn1 * nn2 * nn3).reshape(shape)
    rd, rs = O.rlft3(x.copy(), np.zeros((nn1, 2 * nn2)), 1, mt=True)
    slabs = [torch.from_numpy(np.ascontiguousarray(x[:, r * Y:(r + 1) * Y, :]).ravel()).cuda() for r in range(G)]
    speqs = [torch.zeros(2 * X * nn2, dtype=torch.float64, device="cuda") for _ in range(G)]
    run_direction(gpu, torch, slabs, speqs, shape, G, 1)
    for r in range(G):
        assert cases.rel(slabs[r].cpu().numpy(), rd[r * X:(r + 1) * X]) <= cases.tol(x.size), (r, "data")
        assert cases.rel(speqs[r].cpu().numpy(), rs[r * X:(r + 1) * X]) <= cases.tol(x.size), (r, "speq")
    run_direction(gpu, torch, slabs, speqs, shape, G, -1)
    for r in range(G):
        back = slabs[r].cpu().numpy() * (2.0 / x.size)
        assert cases.rel(back, np.ascontiguousarray(x[:, r * Y:(r + 1) * Y, :])) <= cases.tol(x.size), (r, "round trip")
