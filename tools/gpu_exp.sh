#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W="rlft3_512 four1_12_4096 four1_20_64 fourn2d_8192"
for v in skel skelp; do
  echo "##### $v"; NUMRS_B200_LIB=$PWD/variants/lib_$v.so timeout 300 python tools/kernel_table.py $W 2>&1 | grep -v Traceback
done
