#!/bin/bash
# Last evidence run of round 2 on ONE GPU (≈ 5 GPU-minutes): the whole GPU test tier and smoke on the final library, the default
# bench line, kernel tables of the conv / trig workloads for the shipped library and the twiddle-prefetch variant
# (variants/lib_twpre.so = -DNRB_TW_PREFETCH=1), and that variant's own GPU tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r02_u
timeout 270 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.txt 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${R}_pytest_gpu.txt; tail -3 gpurun_out/${R}_pytest_gpu.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.txt 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${R}_smoke.txt; tail -2 gpurun_out/${R}_smoke.txt
timeout 120 python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench_rlft3_512.json 2> gpurun_out/${R}_bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/${R}_bench_rlft3_512.json
W="convlv_22_16 correl_22_16 autocorrel_22_16 correlnorm_22_16 cosft1_12_4096 cosft2_12_4096 sinft_12_4096 twofft_12_4096"
for v in main twpre; do
  lib=numrs_b200/libnumrs_b200.so; [ $v != main ] && lib=variants/lib_$v.so
  NUMRS_B200_LIB=$PWD/$lib timeout 60 python tools/kernel_table.py $W > gpurun_out/${R}_kernel_table_$v.txt 2>&1
  echo "---- $v"; grep -h "^==\|conv_mid\|aux_normalize" gpurun_out/${R}_kernel_table_$v.txt | cut -c1-120
done
NUMRS_B200_LIB=$PWD/variants/lib_twpre.so timeout 100 python -m pytest tests -m gpu -x -q -k "(trig or twofft or cosft or conv_fused or convlv or correl) and not test_full_size" > gpurun_out/${R}_pytest_gpu_twpre.txt 2>&1; echo "twpre pytest exit $?" | tee -a gpurun_out/${R}_pytest_gpu_twpre.txt; tail -3 gpurun_out/${R}_pytest_gpu_twpre.txt
