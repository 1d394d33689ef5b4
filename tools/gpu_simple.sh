#!/bin/bash
# A/B of the cheap addressing path (simple_addr) through the C ABI
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=tools/cabi_bench
L=numrs_b200/libnumrs_b200.so
run() { echo "== $*"; timeout 25 $B $L "$@" 2>&1 | grep -v "^option" | grep -v "L512 "; }
{
for wl in rlft3:512 four1:12:4096 four1:13:8192 four1:20:64 fourn:8192x8192; do
  run $wl simple_addr=0
  run $wl simple_addr=1
done
} > gpurun_out/simple_ab.txt 2>&1
cat gpurun_out/simple_ab.txt
