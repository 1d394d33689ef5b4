#!/usr/bin/env python
"""torchrun diagnostic: per-phase device times of the slab rlft3 (stage 0 | all-to-all | stage 1)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import numrs_b200 as nb  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.cuda.set_device(lr)
lib = nb.lib()
lib.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from numrs_b200.dist_rlft3 import SlabRlft3  # noqa: E402
mode = sys.argv[2] if len(sys.argv) > 2 else "nccl"
barrier = sys.argv[3] if len(sys.argv) > 3 else "flags"
S = SlabRlft3(lib, n, n, n, mode=mode)
slab = S.plan
f64 = dict(dtype=torch.float64, device="cuda")
buf = torch.empty(slab.local_doubles(), **f64)
speq = torch.empty(slab.speq_doubles(), **f64)
if mode == "nccl":
    send, recv = S.send, S.recv
st = torch.cuda.current_stream().cuda_stream
lib.fill_uniform_device(buf.data_ptr(), 1006, rank * buf.numel(), buf.numel(), st)
acc = [0.0] * 7
reps = 12
flag = torch.zeros(1, dtype=torch.int32, device="cuda")


def direction(isign, e0, e1, e2, e3, call):
    e0.record()
    if mode == "nccl":
        slab.stage(0, isign, buf.data_ptr(), speq.data_ptr(), send.data_ptr(), 0, st)
        e1.record()
        dist.all_to_all_single(recv, send)
        e2.record()
        slab.stage(1, isign, buf.data_ptr(), speq.data_ptr(), 0, recv.data_ptr(), st)
    else:
        slab.set_peers(S._peers[call & 1])
        slab.stage(0, isign, buf.data_ptr(), speq.data_ptr(), 0, 0, st)
        e1.record()
        if barrier == "flags":
            slab.barrier(0, call // 2 + 1, st)
            slab.barrier(1, call // 2 + 1, st)
        else:
            dist.all_reduce(flag)
        e2.record()
        slab.stage(1, isign, buf.data_ptr(), speq.data_ptr(), 0, 0, st)
    e3.record()


ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
call = 0
for it in range(reps + 3):
    dist.barrier()
    torch.cuda.synchronize()
    direction(1, ev[0], ev[1], ev[2], ev[3], call)
    direction(-1, ev[4], ev[5], ev[6], ev[7], call + 1)
    call += 2
    torch.cuda.synchronize()
    buf.mul_(2.0 / n ** 3)
    if it >= 3:
        for j, (a, b) in enumerate([(0, 1), (1, 2), (2, 3), (4, 5), (5, 6), (6, 7)]):
            acc[j] += ev[a].elapsed_time(ev[b])
        acc[6] += ev[0].elapsed_time(ev[7])
t = torch.tensor(acc, device="cuda") / reps
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    xn = "all-to-all" if mode == "nccl" else "barrier (stores already landed)"
    names = ["fwd stage0 (z + x)", "fwd " + xn, "fwd stage1 (y)", "inv stage0 (y)", "inv " + xn, "inv stage1 (x + z)", "total step"]
    a2a_bytes = 8.0 * slab.xchg_doubles() * (world - 1) / world
    print(f"== slab rlft3 {n}^3 on {world} GPUs, exchange={mode} barrier={barrier if mode == "fused" else "-"} (max over ranks, ms)")
    for nm, v in zip(names, t.tolist()):
        extra = f"   {a2a_bytes / v / 1e6:.0f} GB/s per GPU per direction" if "all-to-all" in nm else ""
        print(f"   {v:8.4f}  {nm}{extra}")
S.close()
dist.destroy_process_group()
