#!/usr/bin/env python
"""Worker of tests/test_gpu_multirank.py: one rank of a real multi-process slab transform (launched by
torch.distributed.run).  Every rank transforms its slab with SlabRlft3 (peer stores through CUDA-IPC mappings, epoch-flag
barrier kernels; or NCCL all_to_all_single) and compares the FORWARD SPECTRUM of its nn1-slab element-wise with the oracle
-- not a round trip, which a consistent permutation error in both directions would survive -- then the inverse.

With fewer GPUs than ranks the ranks share devices (rank r -> device r mod ndev): the IPC mapping, the peer-pointer table,
the flag barrier and the exchange addressing are exactly the multi-GPU code path; only the wire is missing.  The process
group is then gloo (NCCL refuses two ranks on one device), which the fused exchange never uses on the data path anyway."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402
import numrs_b200 as nb  # noqa: E402
import oracle as O  # noqa: E402
from numrs_b200.dist_rlft3 import SlabRlft3  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="64x128x32")
    ap.add_argument("--kind", default="rlft3", choices=["rlft3", "fourn"])
    ap.add_argument("--mode", default="fused", choices=["fused", "nccl", "dma"])
    ap.add_argument("--chunks", type=int, default=1)
    ap.add_argument("--reps", type=int, default=3, help="forward + inverse repetitions (exercises the double-buffered receive buffers)")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    ndev = torch.cuda.device_count()
    shared = ndev < world
    dev = local % ndev
    torch.cuda.set_device(dev)
    lib = nb.lib()
    lib.set_device(dev)
    if shared:
        assert a.mode != "nccl", "the NCCL exchange needs one GPU per rank"
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    shape = tuple(int(v) for v in a.shape.split("x"))
    nn1, nn2, nn3 = shape
    G = world
    X, Y = nn1 // G, nn2 // G
    n = nn1 * nn2 * nn3
    real = a.kind == "rlft3"

    def problem(rep):
        """Input slab of this rank and the oracle's forward spectrum for repetition `rep`: DIFFERENT data every time, so a
        stale read of an exchange buffer (they are reused every other call) cannot pass for a fresh one."""
        if real:
            x = O.fill_uniform(1006 + rep, 0, n).reshape(shape)
            rd, rs = O.rlft3(x.copy(), np.zeros((nn1, 2 * nn2)), 1, mt=False)
            return (np.ascontiguousarray(x[:, rank * Y:(rank + 1) * Y, :]).ravel(), np.ascontiguousarray(rd[rank * X:(rank + 1) * X]).ravel(),
                    np.ascontiguousarray(rs[rank * X:(rank + 1) * X]).ravel())
        xf = O.fill_uniform(1008 + rep, 0, 2 * n)
        ref = O.fourn(xf.copy(), list(shape), 1).reshape(nn1, nn2, 2 * nn3)
        xv = xf.reshape(nn1, nn2, 2 * nn3)
        return np.ascontiguousarray(xv[:, rank * Y:(rank + 1) * Y, :]).ravel(), np.ascontiguousarray(ref[rank * X:(rank + 1) * X]).ravel(), None
    scale = (2.0 / n) if real else (1.0 / n)
    slab = SlabRlft3(lib, nn1, nn2, nn3, mode=a.mode, chunks=a.chunks, kind=a.kind)
    s = torch.zeros(slab.speq_doubles, dtype=torch.float64, device="cuda") if real else None
    worst = 0.0
    for rep in range(a.reps):
        mine, want, want_speq = problem(rep)
        assert slab.local_doubles == mine.size
        d = torch.from_numpy(mine.copy()).cuda()
        slab.transform(d, s, 1)
        torch.cuda.synchronize()
        e = cases.rel(d.cpu().numpy(), want)
        if real:
            e = max(e, cases.rel(s.cpu().numpy(), want_speq))
        assert e <= cases.tol(n), f"rank {rank} rep {rep}: forward spectrum differs from the oracle: {e:.3e}"
        worst = max(worst, e)
        if rep % 2 == 1:      # an extra forward call shifts the parity of the double-buffered exchange buffers
            d2 = torch.from_numpy(mine.copy()).cuda()
            slab.transform(d2, s, 1)
            torch.cuda.synchronize()
            assert cases.rel(d2.cpu().numpy(), want) <= cases.tol(n), f"rank {rank} rep {rep}: repeated forward call differs"
        slab.transform(d, s, -1)
        d.mul_(scale)
        torch.cuda.synchronize()
        e = cases.rel(d.cpu().numpy(), mine)
        assert e <= cases.tol(n), f"rank {rank} rep {rep}: inverse differs from the input slab: {e:.3e}"
        worst = max(worst, e)
    slab.close()
    dist.barrier()
    if rank == 0:
        print(f"multirank_slab ok: kind={a.kind} shape={a.shape} ranks={world} devices={ndev} mode={a.mode} chunks={a.chunks} worst rel-L2 {worst:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
