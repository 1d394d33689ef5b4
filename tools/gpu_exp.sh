#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
W="rlft3_512 four1_12_4096 four1_20_64 fourn2d_8192 convlv_22_16 correl_22_16"
timeout 300 python tools/kernel_table.py $W 2>&1 | grep -v Traceback
