#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=tools/cabi_bench
L=numrs_b200/libnumrs_b200.so
run() { lib=$1; shift; echo "== $lib $*"; timeout 60 $B $lib "$@" 2>&1 | grep -v "^option" | grep -v "L512 \|L1024 \|pad_resp\|L4096 "; }
{
for wl in convlv:22:64 correl:22:64 autocorrel:22:16; do
  run $L $wl
  run variants/lib_mid8.so $wl
done
run $L four1:20:64
run $L four1:20:64 tma_col_mask=1536
run $L four1:22:16 tma_col_mask=640
run $L fourn:8192x8192 tma_col_mask=704
} > gpurun_out/r02_i_ab.txt 2>&1
cat gpurun_out/r02_i_ab.txt
