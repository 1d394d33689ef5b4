#!/bin/bash
# TMA-fed strided pass: parity test, then A/B through the C-ABI driver
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tma" 2>&1 | tail -5 | tee gpurun_out/r02_e_tma_tests.txt
B=tools/cabi_bench
L=numrs_b200/libnumrs_b200.so
run() { echo "== $*"; timeout 60 $B $L "$@" 2>&1 | grep -v "^option" | grep -v "L512 \|L1024 "; }
{
run rlft3:512
run rlft3:512 tma_col_mask=512
run rlft3:512 tma_col_mask=512 tma_persist=1
run fourn:512x512x512 tma_col_mask=512
run fourn:8192x8192
run fourn:8192x8192 tma_col_mask=128
run rlft3:1024 tma_col_mask=1024
run rlft3:256
run rlft3:256 tma_col_mask=256
run rlft3:256 tma_col_mask=256 tma_persist=1
} > gpurun_out/r02_e_tma_ab.txt 2>&1
cat gpurun_out/r02_e_tma_ab.txt
