#!/usr/bin/env python
"""torchrun diagnostic: timeline of the DMA-pipelined slab exchange (rank 0, ms since the start of a direction).
Usage: torchrun ... tools/slab_dma_timeline.py <n> <chunks>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import numrs_b200 as nb  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
C = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.cuda.set_device(lr)
lib = nb.lib()
lib.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from numrs_b200.dist_rlft3 import SlabRlft3  # noqa: E402
S = SlabRlft3(lib, n, n, n, mode="dma", chunks=C)
f64 = dict(dtype=torch.float64, device="cuda")
buf = torch.empty(S.local_doubles, **f64)
speq = torch.empty(S.speq_doubles, **f64)
lib.fill_uniform_device(buf.data_ptr(), 1006, rank * buf.numel(), buf.numel(), torch.cuda.current_stream().cuda_stream)
S.plan.dma_timeline(True)
for it in range(6):
    for isign in (1, -1):
        dist.barrier()
        torch.cuda.synchronize()
        S.transform(buf, speq, isign)
        torch.cuda.synchronize()
        text = S.plan.dma_timeline(True)
        if rank == 0 and it == 5:
            print(f"== dma exchange {n}^3, {world} GPUs, chunks={C}, isign={isign}: {text}")
    buf.mul_(2.0 / n ** 3)
S.close()
dist.destroy_process_group()
