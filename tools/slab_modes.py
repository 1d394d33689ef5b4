#!/usr/bin/env python
"""torchrun diagnostic: one process group, many exchange configurations of the slab-decomposed 3-D transforms, timed back
to back on the same box (CUDA events, max over ranks): fused peer stores (1 / 2 / 4 z-chunks), copy-engine DMA (chunks x
copy streams), NCCL all-to-all.  Prints one line per configuration: ms per forward + inverse, whole-job algorithmic GB/s,
speed-up over the given 1-GPU time, NVLink GB/s if the step were all exchange, and the round-trip error.

    torchrun --nproc-per-node 8 tools/slab_modes.py 512 [ms of the 1-GPU step] [kind] [configs ...]
    config = mode[:chunks[:dma_streams[:pull_eighths[:z_chunks]]]]   e.g. fused (push + pull, half and half) fused:1:1:6 push fused:2 dma:2:4 nccl
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import numrs_b200 as nb  # noqa: E402
from numrs_b200.dist_rlft3 import SlabRlft3  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
t1 = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
kind = sys.argv[3] if len(sys.argv) > 3 else "rlft3"
configs = sys.argv[4:] or ["push", "fused", "fused:1:1:3", "fused:1:1:5", "fused:2", "dma:2:4", "nccl"]
torch.cuda.set_device(lr)
lib = nb.lib()
lib.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
real = kind == "rlft3"
vol = n ** 3
bytes_dir = (16.0 * vol + 16.0 * n * n) if real else 32.0 * vol
scale = (2.0 / vol) if real else (1.0 / vol)
f64 = dict(dtype=torch.float64, device="cuda")
steps = 20 if n <= 512 else 6
for cfg in configs:
    parts = cfg.split(":")
    mode = parts[0]
    chunks = int(parts[1]) if len(parts) > 1 else 1
    streams = int(parts[2]) if len(parts) > 2 else 1
    eighths = int(parts[3]) if len(parts) > 3 else 4
    lib.set_option("z_chunks", int(parts[4]) if len(parts) > 4 else 1)
    lib.set_option("dma_streams", streams)
    lib.set_option("pull_eighths", eighths)
    pull = mode != "push"
    if mode == "push":      # the fused exchange with everything pushed by stage 0 (round 1's form)
        mode = "fused"
    try:
        S = SlabRlft3(lib, n, n, n, mode=mode, chunks=chunks, kind=kind, pull=pull)
    except Exception as e:      # noqa: BLE001
        if rank == 0:
            print(f"{cfg:12s} unavailable: {e}", flush=True)
        continue
    pool = [torch.empty(S.local_doubles, **f64) for _ in range(3)]
    speq = torch.empty(S.speq_doubles, **f64) if real else None
    st = torch.cuda.current_stream().cuda_stream
    for b in pool:
        lib.fill_uniform_device(b.data_ptr(), 1006, rank * b.numel(), b.numel(), st)
    ref = pool[0].clone()
    S.transform(pool[0], speq, 1)
    S.transform(pool[0], speq, -1)
    pool[0].mul_(scale)
    t = torch.stack([(pool[0] - ref).pow(2).sum(), ref.pow(2).sum()])
    dist.all_reduce(t)
    err = float(torch.sqrt(t[0] / t[1]))
    for i in range(3):
        S.transform(pool[i % 3], speq, 1)
        S.transform(pool[i % 3], speq, -1)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        S.transform(pool[i % 3], speq, 1)
        S.transform(pool[i % 3], speq, -1)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms[0])
    if rank == 0:
        a2a = S.a2a_bytes_per_gpu()
        print(f"{cfg:12s} {kind} {n}^3 x{world}: {ms:8.3f} ms / step  {2 * bytes_dir / ms / 1e6:8.0f} GB/s" +
              (f"  speed-up {t1 / ms:5.2f}x" if t1 else "") +
              f"  exchange {a2a / 1e6:.1f} MB/GPU/direction = {2 * a2a / ms / 1e6:.0f} GB/s of NVLink if the step were all exchange  round trip {err:.1e}", flush=True)
    S.close()
    del pool, S
    torch.cuda.empty_cache()
dist.barrier()
dist.destroy_process_group()
