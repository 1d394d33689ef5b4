#!/usr/bin/env python
"""Summarise `-Xptxas -v` output (build.log) per FFT pass kernel: registers, stack, spills."""
import re
import sys

txt = open(sys.argv[1] if len(sys.argv) > 1 else "build.log").read()
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
                 r"(\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers")
rows = []
for name, stack, ss, sl, regs in pat.findall(txt):
    m = re.search(r"fft_pass_kernelILi(\d+)ELi(\d+)ELi(n?\d+)ELi(\d+)E", name)
    if m:
        rows.append((int(m.group(1)), "ROW" if m.group(2) == "0" else "COL", "+1" if m.group(3) == "1" else "-1",
                     ["PLAIN", "REAL", "XPOSE"][int(m.group(4))], int(regs), int(stack), int(ss), int(sl)))
    else:
        print(name, "regs", regs, "stack", stack, "spill", ss, sl)
rows.sort()
print("log2n layout dir variant regs stack spill_st spill_ld")
for r in rows:
    print(*r)
