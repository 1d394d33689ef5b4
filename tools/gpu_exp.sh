#!/bin/bash
cd "$(dirname "$0")/.."
for rep in 1 2; do
echo "##### default"; timeout 300 python tools/kernel_table.py rlft3_512 2>&1 | grep -E "real"
echo "##### preidx"; NUMRS_B200_LIB=$PWD/variants/lib_preidx.so timeout 300 python tools/kernel_table.py rlft3_512 2>&1 | grep -E "real"
done
