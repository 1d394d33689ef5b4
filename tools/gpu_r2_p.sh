#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tma" 2>&1 | tail -3 | tee gpurun_out/r02_p_tests.txt
B=tools/cabi_bench
L=numrs_b200/libnumrs_b200.so
run() { echo "== $*"; timeout 60 $B $L "$@" 2>&1 | grep -v "^option" | grep -v "L512 \|pad_resp\|L4096 \|L1 "; }
{
run rlft3:512
run rlft3:512 tma_in_mask=512
run rlft3:512 tma_in_mask=512 tma_in_ctas=3
run convlv:22:64
run convlv:22:64 tma_in_mask=512
run convlv:22:64 tma_in_mask=512 tma_in_ctas=3
run four1:20:64 tma_in_mask=1024
run fourn:8192x8192 tma_in_mask=192
run four1:22:16 tma_in_mask=384
} > gpurun_out/r02_p_tma_in_ab.txt 2>&1
cat gpurun_out/r02_p_tma_in_ab.txt
