#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 10 -c 10 -o gpurun_out/r01_rlft3_full -f python tools/profile_rlft3.py 512 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 2 -c 2 -o gpurun_out/r01_four1_20 -f python tools/profile_generic.py four1_20_64 > gpurun_out/ncu_four1_20.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 1 -c 1 -o gpurun_out/r01_four1_12 -f python tools/profile_generic.py four1_12_4096 > gpurun_out/ncu_four1_12.log 2>&1
tail -2 gpurun_out/ncu_full.log gpurun_out/ncu_four1_20.log gpurun_out/ncu_four1_12.log
ls -la gpurun_out/*.ncu-rep
