#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "convlv or correl" 2>&1 | tail -3
for fl in 1 0; do
echo "##### conv_transposed $fl"; NRB_CONV_TRANSPOSED=$fl timeout 300 python tools/kernel_table.py convlv_22_16 correl_22_16 autocorrel_22_16 convlv_20_64 2>&1
done
