#!/bin/bash
cd "$(dirname "$0")/.."
W="rlft3_512 fourn2d_8192 four1_20_64"
echo "##### default"; timeout 300 python tools/kernel_table.py $W 2>&1 | grep -E "col|^=="
for v in c16 c16m2; do
echo "##### $v"; NUMRS_B200_LIB=$PWD/variants/lib_$v.so timeout 300 python tools/kernel_table.py $W 2>&1 | grep -E "col|^=="
done
