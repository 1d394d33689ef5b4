#!/usr/bin/env python
"""Static SASS evidence for the hottest kernels of libnumrs_b200.so (no GPU needed): writes, per kernel, the instruction
mix by class and the full listing.

    python tools/sass_report.py [lib] [outdir]      -> profiles/r02_sass_<kernel>.txt + profiles/r02_sass_mix.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "numrs_b200", "libnumrs_b200.so")
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles")
HOT = {   # mangled-name fragment -> (file tag, what it is)
    "fft_pass_kernelILi9ELi1ELi1ELi0E": ("col_plain_n512_fwd", "strided 512-point pass: y and x passes of rlft3 / fourn 512^3 (dominant kernel of the headline step)"),
    "fft_pass_kernelILi8ELi0ELi1ELi1E": ("row_real_n256_fwd", "contiguous real 512-point pass, forward: z pass of rlft3 512^3 (untangling in registers by warp shuffles)"),
    "fft_pass_kernelILi8ELi0ELin1ELi1E": ("row_real_n256_inv", "contiguous real 512-point pass, inverse: z pass of rlft3 512^3"),
    "fft_pass_kernelILi10ELi1ELi1ELi2E": ("col_xpose_n1024_fwd", "transposing 1024-point pass: first pass of a 2^20 transform"),
    "fft_pass_kernelILi13ELi0ELi1ELi0E": ("row_plain_n8192_fwd", "contiguous 8192-point pass: fourn 8192^2 rows"),
    "fft_pass_kernelILi12ELi0ELi1ELi0E": ("row_plain_n4096_fwd", "contiguous 4096-point pass: batched four1 4096 x 4096"),
    "conv_mid_kernelILi12E": ("conv_mid_n4096", "fused middle of convlv / correl: forward pass + untangle * spectral op * re-tangle + inverse pass"),
    "fft_col_tma_kernelILi9ELi1ELb0E": ("col_tma_n512_fwd", "TMA-fed strided 512-point pass (experiment, option tma_col_mask): UTMALDG / UTMASTG + mbarrier"),
}
CLASSES = [("fp64", r"^(DFMA|DADD|DMUL|DSETP|DMNMX)"), ("global ld/st", r"^(LDG|STG|LD\b|ST\b|LDGSTS)"), ("shared ld/st", r"^(LDS|STS|LDSM)"),
           ("tma / mbarrier", r"^(UTMALDG|UTMASTG|UTMAPF|UTMACMDFLUSH|UTMACCTL|SYNCS|UBLKCP|FENCE|ELECT)"),
           ("integer / address", r"^(IMAD|IADD|LEA|LOP|SHF|ISETP|SEL|IABS|PRMT|VIADD|UIMAD|UIADD|ULEA|ULOP|USHF|UISETP|USEL|UMOV|MOV|R2UR|S2R|S2UR|CS2R|I2F|F2I|LDC|ULDC)"),
           ("control / sync", r"^(BRA|BAR|WARPSYNC|BSSY|BSYNC|EXIT|CALL|RET|NOP|DEPBAR|MEMBAR|ERRBAR|CCTL|YIELD|NANOSLEEP|PLOP3|UPLOP3|P2R|R2P|VOTE|SHFL|BMSK|UFLO|POPC|BREV|FLO|UPOPC)")]
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", text)
rows = []
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    for frag, (tag, what) in HOT.items():
        if frag in name:
            ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", f)
            mix = collections.Counter()
            for op in ops:
                base = op.split(".")[0]
                for cname, pat in CLASSES:
                    if re.match(pat, base):
                        mix[cname] += 1
                        break
                else:
                    mix["other"] += 1
            total = sum(mix.values())
            with open(os.path.join(out, f"r02_sass_{tag}.txt"), "w") as fh:
                fh.write(f"// {what}\n// cuobjdump -sass numrs_b200/libnumrs_b200.so, function {name}\n// {total} instructions: " +
                         ", ".join(f"{k} {v}" for k, v in mix.most_common()) + "\n" + "Function : " + f)
            rows.append((tag, what, total, mix))
with open(os.path.join(out, "r02_sass_mix.md"), "w") as fh:
    fh.write("# Static SASS instruction mix of the hottest kernels (round 2, `tools/sass_report.py`, full listings in `profiles/r02_sass_*.txt`)\n\n")
    fh.write("| kernel | what | instructions | " + " | ".join(c for c, _ in CLASSES) + " | other |\n|---|---|---:|" + "---:|" * (len(CLASSES) + 1) + "\n")
    for tag, what, total, mix in sorted(rows):
        fh.write(f"| `{tag}` | {what} | {total} | " + " | ".join(f"{mix.get(c, 0)} ({100.0 * mix.get(c, 0) / total:.0f} %)" for c, _ in CLASSES) +
                 f" | {mix.get('other', 0)} |\n")
    fh.write("\nOnly the TMA-fed experiment contains Blackwell / Hopper bulk-copy instructions (`UTMALDG`, `UTMASTG`, `SYNCS`); the shipped default\n"
             "kernels are LDG / STG / LDS / STS + DFMA code: an f64 FFT has no use for tcgen05, and the TMA path measured slower (r02_tuning.md #47).\n")
print(open(os.path.join(out, "r02_sass_mix.md")).read())
