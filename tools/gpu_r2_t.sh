#!/bin/bash
# ncu --set full (with source correlation) of the fused conv middle kernel inside convlv 16 x 2^22
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r02_t
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_mid -s 1 -c 1 -o gpurun_out/${R}_conv_mid_full -f python tools/profile_generic.py convlv_22_16 > gpurun_out/${R}_ncu_mid.log 2>&1
tail -3 gpurun_out/${R}_ncu_mid.log
