#!/usr/bin/env python
"""Small instances of the one-kernel cosft1 / cosft2 / sinft / twofft path (trig_fused.cuh) for compute-sanitizer
memcheck / racecheck: every tile geometry (many lines per CTA, one line per CTA of 256 / 512 / 1024 threads), ragged tiles."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import numrs_b200 as nb  # noqa: E402

L = nb.lib()
for n, cnt in ((16, 130), (64, 33), (512, 5), (2048, 3), (4096, 3), (8192, 2), (16384, 2)):
    cases.check_trig_batch(L, n, cnt)
for n, cnt in ((8, 300), (64, 33), (1024, 3), (2048, 3), (4096, 2), (8192, 2)):
    cases.check_twofft_plan(L, n, cnt)
print("sanitize_trig: all cases passed")
