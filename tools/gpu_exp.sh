#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 ./tools/cluster_exchange_bench 2>&1 | tee gpurun_out/cluster_exchange_bench.log
timeout 600 python -m pytest tests -m gpu -x -q -k "cosft or correl_normalized or device_resident or host_mirror" 2>&1 | tail -3
timeout 300 python tools/kernel_table.py cosft1_22_16 cosft1_12_4096 cosft2_22_16 cosft2_12_4096 sinft_12_4096 correlnorm_22_16 correlnormfast_22_16 fourn2d_8192 2>&1 | tee gpurun_out/r01_kernel_table_next2.txt
