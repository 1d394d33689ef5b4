#!/bin/bash
# One-kernel cosine / sine / twofft path, second version (batched loads, rotated twiddles, shuffle-assisted stores): sanitizer,
# its GPU tests, kernel tables of the tile-size / occupancy variants (tools: variants/lib_trig_*.so), one ncu capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r02_s
( time timeout 180 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_trig.py ) > gpurun_out/${R}_sanitizer_racecheck.txt 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/${R}_sanitizer_racecheck.txt
( time timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_trig.py ) > gpurun_out/${R}_sanitizer_memcheck.txt 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/${R}_sanitizer_memcheck.txt
grep -h "SUMMARY" gpurun_out/${R}_sanitizer_*.txt
timeout 400 python -m pytest tests -m gpu -x -q -k "trig or twofft or cosft or sinft or next_rows or conv_fused or golden or device_resident" > gpurun_out/${R}_pytest_gpu.txt 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${R}_pytest_gpu.txt; tail -3 gpurun_out/${R}_pytest_gpu.txt
W="twofft_12_4096 cosft1_12_4096 cosft2_12_4096 sinft_12_4096 cosft1_13_2048 cosft1_14_1024 cosft1_8_65536 twofft_8_65536"
for v in main t10 t12 m3; do
  lib=numrs_b200/libnumrs_b200.so; [ $v != main ] && lib=variants/lib_trig_$v.so
  NUMRS_B200_LIB=$PWD/$lib timeout 200 python tools/kernel_table.py $W > gpurun_out/${R}_kernel_table_$v.txt 2>&1
  echo "---- $v"; grep -h "^==" gpurun_out/${R}_kernel_table_$v.txt | cut -c1-110
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:trig_kernel -s 1 -c 1 -o gpurun_out/${R}_trig_cosft1_4096_full -f python tools/profile_generic.py cosft1_12_4096 > gpurun_out/${R}_ncu_trig.log 2>&1
