#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/kernel_table.py twofft_20_16 twofft_12_4096 correlnorm_22_16 correlnormfast_22_16 autocorrel_22_16 cosft1_22_16 cosft1_12_4096 cosft2_22_16 cosft2_12_4096 sinft_12_4096 > gpurun_out/r01_kernel_table_next.txt 2>&1
cat gpurun_out/r01_kernel_table_next.txt
