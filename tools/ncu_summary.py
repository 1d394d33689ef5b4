#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) per kernel: duration, DRAM bytes, throughput %, occupancy,
registers, L1/smem pipe, top stall reasons.  Usage: python tools/ncu_summary.py file.ncu-rep [out.md]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def g(r, name, default=""):
    return r[col[name]] if name in col else default


def f(r, name):
    try:
        return float(g(r, name).replace(",", ""))
    except ValueError:
        return float("nan")


stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
out = []
out.append(f"# ncu --set full summary: {rep}\n")
out.append("| kernel | grid | block | regs | dur us | dram rd MB | dram wr MB | dram % peak | L1/TEX % | LTS % | fp64 pipe % | issue active % | warps active % | smem conflicts/wavefronts | top stalls (warps per issue) |")
out.append("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in data:
    name = g(r, "Kernel Name").replace("void ", "").replace("(PassParams, unsigned int)", "").replace("nrb::", "")
    stalls = sorted(((f(r, h), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall_cols), reverse=True)[:4]
    conf = f(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
    wav = f(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    out.append("| {} | {} | {} | {} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.0f}/{:.0f} | {} |".format(
        name, g(r, "Grid Size"), g(r, "Block Size"), g(r, "launch__registers_per_thread"),
        f(r, "gpu__time_duration.sum"), f(r, "dram__bytes_read.sum") * (1e3 if "Gbyte" in rows[1][col["dram__bytes_read.sum"]] else 1),
        f(r, "dram__bytes_write.sum") * (1e3 if "Gbyte" in rows[1][col["dram__bytes_write.sum"]] else 1),
        f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), f(r, "l1tex__throughput.avg.pct_of_peak_sustained_active"),
        f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"), f(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        conf, wav, ", ".join(f"{n} {v:.1f}" for v, n in stalls)))
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
