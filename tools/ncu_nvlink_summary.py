#!/usr/bin/env python
"""Summarise the csv of the ncu NVLink pass (tools/profile_multi_nvlink.py): per kernel launch the device, duration, bytes sent
and received over NVLink and the rate they make.  Usage: python tools/ncu_nvlink_summary.py launches.csv [out.md]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
col = {h: i for i, h in enumerate(H)}
launches = {}
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    key = r[col["ID"]]
    d = launches.setdefault(key, {"name": r[col["Kernel Name"]], "dev": r[col["Device"]] if "Device" in col else "?"})
    try:
        v = float(r[col["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[col["Metric Unit"]]
    m = r[col["Metric Name"]]
    if m.startswith("gpu__time_duration"):
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)           # -> us
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)   # -> bytes
    d[m] = v
out = ["| # | device | kernel | us | NVLink sent MB | received MB | sent GB/s | received GB/s |", "|---|---|---|---:|---:|---:|---:|---:|"]
for k, d in launches.items():
    us = d.get("gpu__time_duration.sum", 0.0)
    tx, rx = d.get("nvltx__bytes_data_user.sum", 0.0), d.get("nvlrx__bytes_data_user.sum", 0.0)
    name = d["name"].replace("void nrb::", "").replace("nrb::", "")[:70]
    out.append(f"| {k} | {d['dev']} | `{name}` | {us:.1f} | {tx / 1e6:.1f} | {rx / 1e6:.1f} | {tx / us / 1e3 if us else 0:.0f} | {rx / us / 1e3 if us else 0:.0f} |")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
