#!/bin/bash
# 1-GPU call: TMA variants, conv rest=11 A/B, ncu --set full captures (raw CSV pages kept, .ncu-rep too when small)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=tools/cabi_bench
L=numrs_b200/libnumrs_b200.so
run() { lib=$1; shift; echo "== $lib $*"; timeout 60 $B $lib "$@" 2>&1 | grep -v "^option" | grep -v "L512 \|L1024 \|pad_resp\|L4096 "; }
{
run $L rlft3:512 tma_col_mask=512
run variants/lib_tma3.so rlft3:512 tma_col_mask=512
run variants/lib_tma3.so fourn:512x512x512 tma_col_mask=512
for wl in convlv:22:64 correl:22:64 autocorrel:22:16; do
  run $L $wl
  run $L $wl conv_rest_log2=11
done
} > gpurun_out/r02_f_ab.txt 2>&1
cat gpurun_out/r02_f_ab.txt
# ncu: the ten launches of one rlft3 512^3 forward + inverse (plain kernels), then the TMA-fed variant's strided passes
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:fft_pass_kernel -c 10 -o gpurun_out/r02_rlft3_512_full -f $B $L rlft3:512 > gpurun_out/ncu_f1.log 2>&1
timeout 600 $NCU -k regex:fft_col_tma_kernel -c 2 -o gpurun_out/r02_rlft3_512_tma_full -f $B $L rlft3:512 tma_col_mask=512 > gpurun_out/ncu_f2.log 2>&1
timeout 600 $NCU -k regex:"fft_pass_kernel|conv_mid" -c 5 -o gpurun_out/r02_convlv_22_full -f $B $L convlv:22:16 > gpurun_out/ncu_f3.log 2>&1
timeout 600 $NCU -k regex:fft_pass_kernel -c 2 -o gpurun_out/r02_four1_20_full -f $B $L four1:20:64 > gpurun_out/ncu_f4.log 2>&1
for r in r02_rlft3_512_full r02_rlft3_512_tma_full r02_convlv_22_full r02_four1_20_full; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/$r.ncu-rep gpurun_out/$r.md > /dev/null 2>&1
  ls -la gpurun_out/$r.*
done
# launch list of the bench command (share of the step per kernel)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
# keep the reports only if they fit the 64 MiB return budget
du -sm gpurun_out | tail -1
for r in r02_convlv_22_full r02_four1_20_full; do rm -f gpurun_out/$r.ncu-rep; done
du -sm gpurun_out | tail -1
