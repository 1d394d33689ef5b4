#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tma" 2>&1 | tail -3 | tee gpurun_out/r02_o_tests.txt
B=tools/cabi_bench
L=numrs_b200/libnumrs_b200.so
run() { echo "== $*"; timeout 60 $B $L "$@" 2>&1 | grep -v "^option"; }
{ run four1:20:64; run four1:20:64 tma_xpose=1; run four1:20:64; run four1:20:64 tma_xpose=1; } > gpurun_out/r02_o_xpose_tma_ab.txt 2>&1
cat gpurun_out/r02_o_xpose_tma_ab.txt
