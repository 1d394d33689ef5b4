#!/bin/bash
cd "$(dirname "$0")/.."
W="rlft3_512 four1_8_65536 four1_10_16384 four1_12_4096 fourn2d_8192"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
echo "##### default (twiddle powers)"; timeout 300 python tools/kernel_table.py $W 2>&1 | grep -E "row|^=="
echo "##### table loads"; NUMRS_B200_LIB=$PWD/variants/lib_notwp.so timeout 300 python tools/kernel_table.py $W 2>&1 | grep -E "row|^=="
done
