#!/bin/bash
# NVLink counters of the slab exchange kernels (2 GPUs, ONE process: the multi-device host-slice path), ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r02_w
timeout 75 ncu --metrics nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,gpu__time_duration.sum --clock-control none -k regex:fft_ -c 60 --csv --log-file gpurun_out/${R}_nvlink_launches.csv python tools/profile_multi_nvlink.py 512 > gpurun_out/${R}_nvlink.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/${R}_nvlink.log; python tools/ncu_nvlink_summary.py gpurun_out/${R}_nvlink_launches.csv gpurun_out/${R}_nvlink_summary.md | tail -40
