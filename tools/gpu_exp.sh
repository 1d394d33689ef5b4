#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W="rlft3_512 four1_12_4096 four1_20_64 fourn2d_8192"
for d in 0 148 296 592 1184; do
echo "##### prefetch_dist $d"; NRB_PREFETCH_DIST=$d timeout 300 python tools/kernel_table.py $W 2>&1 | grep -E "^==|_L131072|_L262144|L4096|L65536|L8192|L524288|L1048576"
done
