#!/bin/bash
# First GPU call of the next round: puts the switches that were prepared without GPU time on
# hardware.  C-ABI driver first (seconds), then the gated GPU tests (imports torch: about a minute more).
#   /usr/local/graft/bin/gpurun --timeout 600 -- ./tools/gpu_next.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
[ -x tools/cabi_bench ] || g++ -O2 -std=c++17 -o tools/cabi_bench tools/cabi_bench.cpp -ldl
B=tools/cabi_bench
L=numrs_b200/libnumrs_b200.so
run() { echo "== $*"; timeout 60 $B $L "$@" 2>&1 | grep -v "^option"; }
{
run rlft3:512 speq_side=0
run rlft3:512 speq_side=1
for wl in convlv:22:64 correl:22:64 autocorrel:22:16; do
  run $wl conv_fused_mid=0
  run $wl conv_fused_mid=1      # same checksum line as the run above = same results
done
# library variants (tools/build_variants.sh, run it here before the call): same workloads, other build switches
for v in tw3 xpose simple; do
  V=variants/lib_$v.so
  [ -f $V ] || continue
  for wl in rlft3:512 four1:12:16384 four1:13:8192 four1:20:64 fourn:8192x8192; do
    echo "== $V $wl"; timeout 60 $B $V $wl 2>&1 | grep -v "^option" | grep -v "L512 "
  done
done
} > gpurun_out/next_ab.txt 2>&1
cat gpurun_out/next_ab.txt
NRB_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "side_lane or conv_fused" 2>&1 | tail -5 | tee gpurun_out/next_tests.txt
