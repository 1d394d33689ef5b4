#!/bin/bash
cd "$(dirname "$0")/.."
W="rlft3_512 four1_8_65536 four1_10_16384 four1_11_8192 four1_12_4096 four1_13_2048"
echo "##### default"; timeout 300 python tools/kernel_table.py $W 2>&1 | grep -E "row|^=="
for v in r8all tl9 tl11 split11; do
echo "##### $v"; NUMRS_B200_LIB=$PWD/variants/lib_$v.so timeout 300 python tools/kernel_table.py $W 2>&1 | grep -E "row|^=="
done
