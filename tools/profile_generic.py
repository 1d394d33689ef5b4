#!/usr/bin/env python
"""Short driver for ncu: run a kernel_table workload twice (warm-up + measured) without event profiling."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import numrs_b200 as nb  # noqa: E402

lib = nb.lib()
st = torch.cuda.current_stream().cuda_stream
f64 = dict(dtype=torch.float64, device="cuda")
wl = sys.argv[1]
if wl.startswith("four1_"):
    _, lg, cnt = wl.split("_")
    nn, cnt = 1 << int(lg), int(cnt)
    plan = lib.plan_create(nb.KIND_FOUR1, [nn], batch=cnt)
    buf = torch.empty(2 * nn * cnt, **f64)
    lib.fill_uniform_device(buf.data_ptr(), 1002, 0, buf.numel(), st)
    for _ in range(2):
        plan.exec(buf.data_ptr(), isign=1, stream=st)
elif wl.startswith("convlv_"):
    _, lg, cnt = wl.split("_")
    n, cnt, m = 1 << int(lg), int(cnt), 4096
    a = torch.empty(n * cnt, **f64)
    lib.fill_uniform_device(a.data_ptr(), 1004, 0, a.numel(), st)
    o = torch.empty(n * cnt, **f64)
    aux = torch.empty(m, **f64)
    lib.fill_uniform_device(aux.data_ptr(), 1005, 0, m, st)
    plan = lib.plan_create(nb.KIND_CONVLV, [n, m], batch=cnt)
    for _ in range(2):
        plan.exec(a.data_ptr(), aux.data_ptr(), o.data_ptr(), isign=1, stream=st)
elif wl.split("_")[0] in ("cosft1", "cosft2", "sinft", "twofft"):
    kind, lg, cnt = wl.split("_")
    n, cnt = 1 << int(lg), int(cnt)
    a = torch.empty((n + 2) * cnt, **f64)
    b = torch.empty(n * cnt, **f64)
    lib.fill_uniform_device(a.data_ptr(), 1010, 0, a.numel(), st)
    lib.fill_uniform_device(b.data_ptr(), 1011, 0, b.numel(), st)
    if kind == "twofft":
        o = torch.empty(2 * (2 * n + 2) * cnt, **f64)
        plan = lib.plan_create(nb.KIND_TWOFFT, [n], batch=cnt)
        for _ in range(2):
            plan.exec(a.data_ptr(), b.data_ptr(), o.data_ptr(), isign=1, stream=st)
    else:
        plan = lib.plan_create({"cosft1": nb.KIND_COSFT1, "cosft2": nb.KIND_COSFT2, "sinft": nb.KIND_SINFT}[kind], [n], batch=cnt)
        for _ in range(2):
            plan.exec(a.data_ptr(), isign=1, stream=st)
torch.cuda.synchronize()
print("launches:", plan.num_launches(1))
