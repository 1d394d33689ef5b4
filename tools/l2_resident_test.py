#!/usr/bin/env python
"""How fast do the FFT pass kernels run when their working set stays in L2?  Repeats a plan on ONE
buffer of a given size and reports per-kernel algorithmic GB/s vs buffer size."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import numrs_b200 as nb  # noqa: E402

lib = nb.lib()
st = torch.cuda.current_stream().cuda_stream
f64 = dict(dtype=torch.float64, device="cuda")
print("buffer MiB | row n4096 (four1 batch) | col n512 + row n256 (fourn [512, X, 256] axes 0 and 2)")
for mib in (8, 16, 32, 48, 64, 96, 128, 256, 1024):
    elems = mib * (1 << 20) // 16
    # ROW kernel: batch of 4096-point transforms
    cnt = elems // 4096
    plan = lib.plan_create(nb.KIND_FOUR1, [4096], batch=cnt)
    buf = torch.zeros(2 * elems, **f64)
    for _ in range(3):
        plan.exec(buf.data_ptr(), isign=1, stream=st)
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.exec(buf.data_ptr(), isign=1, stream=st)
    e1.record()
    torch.cuda.synchronize()
    row = 32.0 * elems * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    plan.destroy()
    # COL kernel n512 on [512][X][256]: fourn over 3 dims; profile to split kernels
    X = elems // (512 * 256)
    out = ""
    if X >= 2:
        plan = lib.plan_create(nb.KIND_FOURN, [512, X, 256], batch=1)
        for _ in range(3):
            plan.exec(buf.data_ptr(), isign=1, stream=st)
        agg = {}
        for _ in range(8):
            for name, b, ms in plan.profile(buf.data_ptr(), isign=1, stream=st):
                a = agg.setdefault(name, [0.0, 0.0])
                a[0] += b
                a[1] += ms
        out = "  ".join(f"{k}: {v[0] / v[1] / 1e6:6.0f}" for k, v in agg.items())
        plan.destroy()
    print(f"{mib:6d} MiB | {row:7.0f} GB/s | {out}")
    del buf
