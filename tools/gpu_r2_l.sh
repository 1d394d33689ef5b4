#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=tools/cabi_bench
run() { lib=$1; shift; echo "== $lib $*"; timeout 60 $B $lib "$@" 2>&1 | grep -v "^option" | grep -v "L512 \|L1024 \|pad_resp\|L4096 "; }
{
for l in numrs_b200/libnumrs_b200.so variants/lib_notwrot.so numrs_b200/libnumrs_b200.so variants/lib_notwrot.so; do run $l rlft3:512; done
for l in numrs_b200/libnumrs_b200.so variants/lib_notwrot.so; do run $l convlv:13:65536; run $l rlft3:1024; done
} > gpurun_out/r02_l_twrot_ab.txt 2>&1
grep "row_real\|^==\|forward +\|one call" gpurun_out/r02_l_twrot_ab.txt
