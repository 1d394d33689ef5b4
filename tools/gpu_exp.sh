#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for lag in 4 8 16 32 64; do echo "### lag $lag"; NRB_FUSE_LAG=$lag timeout 120 python tools/kernel_table.py rlft3_512 2>&1 | grep -v Traceback; done
echo "### unfused"; NRB_FUSE_ZY=0 timeout 120 python tools/kernel_table.py rlft3_512 2>&1 | grep -v Traceback
