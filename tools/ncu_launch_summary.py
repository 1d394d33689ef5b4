#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel and grid size.
Usage: python tools/ncu_launch_summary.py launches.csv out.md "<command that was profiled>" """
import csv
import re
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, data = rows[0], rows[1:]
col = {h: i for i, h in enumerate(hdr)}
agg, order = {}, []
for r in data:
    name = r[col["Kernel Name"]]
    m = re.search(r"fft_pass_kernel<(?:\(int\))?(-?\d+), (?:\(int\))?(-?\d+), (?:\(int\))?(-?\d+), (?:\(int\))?(-?\d+)>", name)
    if m:
        lg, lay, d, var = (int(x) for x in m.groups())
        name = f"fft_{'row' if lay == 0 else 'col'}_{['plain', 'real', 'xpose'][var]}_n{1 << lg}_{'p' if d > 0 else 'm'}"
    m = re.search(r"fft_col_tma_kernel<(?:\(int\))?(-?\d+), (?:\(int\))?(-?\d+)", name)
    if m:
        name = f"fft_col_tma_n{1 << int(m.group(1))}_{'p' if int(m.group(2)) > 0 else 'm'}"
    key = f"{name[:60]} grid {r[col['Grid Size']]}"
    if key not in agg:
        agg[key] = [0, 0.0, name.startswith("fft_")]
        order.append(key)
    agg[key][0] += 1
    agg[key][1] += float(r[col["Metric Value"]]) / 1e3
fft_total = sum(v[1] for v in agg.values() if v[2])
out = [f"# ncu launch list (gpu__time_duration.sum, --clock-control none): `{cmd}`\n",
       "Cold-cache, serialised per-launch times: compare SHARES with bench.py's event timings, not absolutes.\n",
       "| kernel (grid) | launches | total us | avg us | share of FFT time |", "|---|---:|---:|---:|---:|"]
for k in order:
    n, us, is_fft = agg[k]
    out.append(f"| {k} | {n} | {us:.1f} | {us / n:.2f} | {100 * us / fft_total if is_fft else 0.0:.1f} % |")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
