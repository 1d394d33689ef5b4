#!/bin/bash
# 2-GPU call: exchange micro-benchmark + 8192-point COL tile variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 tools/p2p_copy_bench 256 > gpurun_out/r02_c_p2p_copy_bench.txt 2>&1
timeout 60 tools/p2p_copy_bench 32 >> gpurun_out/r02_c_p2p_copy_bench.txt 2>&1
cat gpurun_out/r02_c_p2p_copy_bench.txt
B=tools/cabi_bench
{
for v in numrs_b200/libnumrs_b200.so variants/lib_tlA.so variants/lib_tlD.so; do for wl in four1:20:64 rlft3:1024; do echo "== $v $wl"; timeout 60 $B $v $wl 2>&1 | grep -v "^option" | grep -v "L1024 \|L512 "; done; done
for v in numrs_b200/libnumrs_b200.so variants/lib_tlB.so variants/lib_tlC.so; do for wl in rlft3:512 fourn:512x512x512; do echo "== $v $wl"; timeout 60 $B $v $wl 2>&1 | grep -v "^option" | grep -v "L1024 \|L512 "; done; done
} > gpurun_out/r02_c_tile13_ab.txt 2>&1
cat gpurun_out/r02_c_tile13_ab.txt
