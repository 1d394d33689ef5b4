#!/bin/bash
# Round-2 factorisation A/Bs through planner options (C-ABI driver, seconds each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
[ -x tools/cabi_bench ] || g++ -O2 -std=c++17 -o tools/cabi_bench tools/cabi_bench.cpp -ldl
B=tools/cabi_bench
L=numrs_b200/libnumrs_b200.so
run() { echo "== $*"; timeout 60 $B $L "$@" 2>&1 | grep -v "^option"; }
{
run four1:20:64
run four1:20:64 col_max_log2=7
run four1:20:64 col_max_log2=9
run four1:13:8192
run four1:13:8192 row_max_log2=12
run fourn:8192x8192
run fourn:8192x8192 row_max_log2=12
run four1:12:16384
run four1:12:16384 row_max_log2=11
run four1:22:16
run four1:22:16 col_max_log2=8
run fourn:512x512x512
run rlft3:1024
} > gpurun_out/r02_b_factor_ab.txt 2>&1
cat gpurun_out/r02_b_factor_ab.txt
