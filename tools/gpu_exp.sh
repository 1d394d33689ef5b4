#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/kernel_table.py rlft3_512 rlft3_256 four1_12_4096 convlv_22_16 2>&1 | grep -v Traceback
