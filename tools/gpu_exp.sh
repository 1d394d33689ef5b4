#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -2 gpurun_out/sanitizer_racecheck.log
W="rlft3_512 four1_9_32768 four1_10_16384 four1_11_8192 four1_12_4096 four1_13_2048 fourn2d_8192"
echo "##### default (sub mode)"; timeout 300 python tools/kernel_table.py $W 2>&1 | grep -E "row|^=="
echo "##### nosub"; NUMRS_B200_LIB=$PWD/variants/lib_nosub.so timeout 300 python tools/kernel_table.py $W 2>&1 | grep -E "row|^=="
