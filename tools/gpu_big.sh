#!/bin/bash
# A/B of the big-tile passes (fft_pass2.cuh) against the plain passes, C ABI only (tools/cabi_bench.cpp).
# Build first:  g++ -O2 -std=c++17 -o tools/cabi_bench tools/cabi_bench.cpp -ldl
# Variants (profiles/r01_big_tile_ab_2.txt, _3.txt) are library builds with other switches, e.g.
#   make -C numrs_b200/csrc OBJDIR=build_v2ppt8 OUT=../../variants/lib_v2ppt8.so EXTRA="-DNRB_V2_PPT=8"     (8 points/thread)
#   make -C numrs_b200/csrc OBJDIR=build_v2direct OUT=../../variants/lib_v2direct.so EXTRA="-DNRB_V2_DIRECT=1" (no landing zone)
# and are run by passing their path instead of $L.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=numrs_b200/libnumrs_b200.so
B=tools/cabi_bench
run() { echo "== $*"; timeout 25 $B $L "$@" 2>&1 | grep -v "^option"; }
{
run four1:13:8192
run four1:13:8192 big_row_mask=8192
run four1:12:16384
run four1:12:16384 big_row_mask=4096
run four1:11:32768
run four1:11:32768 big_row_mask=2048
run four1:20:64
run four1:20:64 big_col_mask=1024
run fourn:8192x8192
run fourn:8192x8192 big_row_mask=8192
run rlft3:512
run rlft3:512 big_col_mask=512
} > gpurun_out/big_ab.txt 2>&1
cat gpurun_out/big_ab.txt
