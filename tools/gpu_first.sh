#!/bin/bash
# First GPU trip: smoke, parity tests, bench, kernel tables for library variants, ncu launch list + full capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 300 python tools/kernel_table.py rlft3_512 four1_12_4096 four1_20_64 fourn2d_8192 convlv_22_16 correl_22_16 > gpurun_out/ktable_default.log 2>&1
for v in mb3 mb2 r16; do
  NUMRS_B200_LIB=$PWD/variants/lib_$v.so timeout 300 python tools/kernel_table.py rlft3_512 four1_12_4096 four1_20_64 > gpurun_out/ktable_$v.log 2>&1
done
for c in 9 11; do
  NRB_COL_MAX_LOG2=$c timeout 200 python tools/kernel_table.py four1_20_64 fourn2d_8192 convlv_22_16 > gpurun_out/ktable_col$c.log 2>&1
done
for g in 8 16 64; do
  NRB_L2_GROUP_MB=$g timeout 200 python tools/kernel_table.py rlft3_512 > gpurun_out/ktable_l2g$g.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_rlft3.csv python tools/profile_rlft3.py 512 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 62 -c 10 -o gpurun_out/prof_rlft3 -f python tools/profile_rlft3.py 512 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench.err
cat gpurun_out/bench.json | head -c 3000
cat gpurun_out/ktable_default.log
