#!/usr/bin/env python
"""ONE process, two (or more) GPUs: rlft3 on a host volume through the multi-device host-slice path (multi.cpp), forward and
inverse, for an ncu pass that reads the NVLink counters of the exchange kernels:

    ncu --metrics nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,gpu__time_duration.sum --clock-control none \
        -k regex:fft_ --csv --log-file out.csv python tools/profile_multi_nvlink.py [n]

ncu serialises the launches, so every kernel is alone on the link: the numbers are what one push (stage-0 stores into the
peer's receive buffer) and one pull (stage-1 loads from the peer's send buffer) reach uncontended."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import numrs_b200 as nb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
L = nb.lib()
L.set_option("num_devices", 0)
assert L.num_devices_in_use() >= 2, "needs at least two GPUs"
rng = np.random.default_rng(7)
data = rng.uniform(-1, 1, (n, n, n))
orig = data.copy()
speq = np.zeros((n, 2 * n))
before = L.multi_device_calls(0)
nb.rlft3(data, speq, n, n, n, 1)
nb.rlft3(data, speq, n, n, n, -1)
assert L.multi_device_calls(0) - before == 2
err = float(np.linalg.norm((data * (2.0 / n ** 3) - orig).ravel()) / np.linalg.norm(orig.ravel()))
print(f"rlft3 {n}^3 over {L.num_devices_in_use()} GPUs in one process: round-trip rel. L2 {err:.2e}")
assert err < 1e-12
