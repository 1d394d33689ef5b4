#!/usr/bin/env python
"""torchrun diagnostic: timeline of the pipelined slab exchange (events on both streams, ms since the start of the
direction, rank 0).  Usage: torchrun ... tools/slab_timeline.py <n> <chunks>"""
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
S = SlabRlft3(lib, n, n, n, mode="fused", chunks=C)
plan = S.plan
f64 = dict(dtype=torch.float64, device="cuda")
buf = torch.empty(plan.local_doubles(), **f64)
speq = torch.empty(plan.speq_doubles(), **f64)
main = torch.cuda.current_stream()
side = S._side
st, sd = main.cuda_stream, side.cuda_stream
lib.fill_uniform_device(buf.data_ptr(), 1006, rank * buf.numel(), buf.numel(), st)
reps = 10
names = ["pre"] + [f"s0[{c}]" for c in range(C)] + [f"wait[{c}]" for c in range(C)] + [f"s1[{c}]" for c in range(C)] + ["post"]
acc = {isign: [0.0] * len(names) for isign in (1, -1)}


def E():
    return torch.cuda.Event(enable_timing=True)


call = 0
for it in range(reps + 3):
    for isign in (1, -1):
        dist.barrier()
        torch.cuda.synchronize()
        plan.set_peers(S._peers[call & 1])
        epoch = call // 2 + 1
        call += 1
        d, q = buf.data_ptr(), speq.data_ptr()
        e0 = E(); e0.record(main)
        plan.stage_part(0, -1, isign, d, q, st)
        e_pre = E(); e_pre.record(main)
        go = torch.cuda.Event(); go.record(main); side.wait_event(go)
        e_s0, e_w, e_s1 = [], [], []
        for c in range(C):
            plan.stage_part(0, c, isign, d, q, st)
            plan.barrier_chunk(0, c, epoch, st)
            e = E(); e.record(main); e_s0.append(e)
            plan.barrier_chunk(1, c, epoch, sd)
            e = E(); e.record(side); e_w.append(e)
            plan.stage_part(1, c, isign, d, q, sd)
            e = E(); e.record(side); e_s1.append(e)
        done = torch.cuda.Event(); done.record(side); main.wait_event(done)
        plan.stage_part(0, C, isign, d, q, st)
        e_post = E(); e_post.record(main)
        torch.cuda.synchronize()
        if it >= 3:
            evs = [e_pre] + e_s0 + e_w + e_s1 + [e_post]
            for j, e in enumerate(evs):
                acc[isign][j] += e0.elapsed_time(e)
    buf.mul_(2.0 / n ** 3)
if rank == 0:
    print(f"== pipelined slab rlft3 {n}^3, {world} GPUs, chunks={C}, cap={os.environ.get('NRB_XCHG_GRID_CAP', '0')}: event times, ms since start (rank 0)")
    for isign in (1, -1):
        print("  isign", isign, "  ".join(f"{nm}={v / reps:.3f}" for nm, v in zip(names, acc[isign])))
S.close()
dist.destroy_process_group()
