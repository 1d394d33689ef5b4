#!/usr/bin/env python
"""Short driver for ncu: one warm-up and one measured forward+inverse rlft3 (default 512^3)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import numrs_b200 as nb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lib = nb.lib()
st = torch.cuda.current_stream().cuda_stream
plan = lib.plan_create(nb.KIND_RLFT3, [n, n, n])
bufs = [torch.empty(n ** 3, dtype=torch.float64, device="cuda") for _ in range(2)]
sp = torch.empty(2 * n * n, dtype=torch.float64, device="cuda")
for b in bufs:
    lib.fill_uniform_device(b.data_ptr(), 1006, 0, b.numel(), st)
for b in bufs:
    plan.exec(b.data_ptr(), sp.data_ptr(), isign=1, stream=st)
    plan.exec(b.data_ptr(), sp.data_ptr(), isign=-1, stream=st)
torch.cuda.synchronize()
print("launches per direction:", plan.num_launches(1), plan.num_launches(-1))
