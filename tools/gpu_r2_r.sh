#!/bin/bash
# Round-2 evidence run (ONE GPU, ~3 min): the GPU tests that touch the one-kernel cosine / sine / twofft path, its
# compute-sanitizer memcheck + racecheck, the next-row kernel tables with the path on and off, one ncu --set full capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r02_r
( time timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_trig.py ) > gpurun_out/${R}_sanitizer_memcheck.txt 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/${R}_sanitizer_memcheck.txt
( time timeout 180 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_trig.py ) > gpurun_out/${R}_sanitizer_racecheck.txt 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/${R}_sanitizer_racecheck.txt
grep -h "SUMMARY" gpurun_out/${R}_sanitizer_*.txt
timeout 400 python -m pytest tests -m gpu -x -q -k "trig or twofft or cosft or sinft or next_rows or conv_fused or golden or device_resident" > gpurun_out/${R}_pytest_gpu.txt 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${R}_pytest_gpu.txt; tail -4 gpurun_out/${R}_pytest_gpu.txt
W="twofft_12_4096 twofft_13_2048 cosft1_12_4096 cosft2_12_4096 sinft_12_4096 cosft1_13_2048 cosft1_14_1024 cosft1_8_65536 sinft_10_16384"
timeout 200 python tools/kernel_table.py $W > gpurun_out/${R}_kernel_table_next.txt 2>&1
NRB_TRIG_FUSED=0 timeout 200 python tools/kernel_table.py $W > gpurun_out/${R}_kernel_table_next_unfused.txt 2>&1
grep -h "^==" gpurun_out/${R}_kernel_table_next.txt gpurun_out/${R}_kernel_table_next_unfused.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:trig_kernel -s 1 -c 1 -o gpurun_out/${R}_trig_cosft1_4096_full -f python tools/profile_generic.py cosft1_12_4096 > gpurun_out/${R}_ncu_trig.log 2>&1
ncu -i gpurun_out/${R}_trig_cosft1_4096_full.ncu-rep --page raw --csv > gpurun_out/${R}_trig_cosft1_4096_full.raw.csv 2>/dev/null
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.txt 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${R}_smoke.txt
