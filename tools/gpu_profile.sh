#!/bin/bash
# Evidence run: GPU tests, bench (ours + reference arm), other workloads, ncu launch list of the bench command,
# ncu --set full of the hot kernels.  Usage: bash tools/gpu_profile.sh r01
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r01}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench_rlft3_512.json 2> gpurun_out/bench_ours.err; echo "bench exit $?" >> gpurun_out/bench_ours.err
for w in four1_batch four1_1m fourn2d convlv correl; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/${R}_bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 300 python tools/kernel_table.py rlft3_512 four1_12_4096 four1_20_64 fourn2d_8192 convlv_22_16 correl_22_16 > gpurun_out/${R}_kernel_table.txt 2>&1
timeout 300 python tools/kernel_table.py twofft_20_16 twofft_12_4096 correlnorm_22_16 correlnormfast_22_16 autocorrel_22_16 cosft1_22_16 cosft1_12_4096 cosft2_22_16 cosft2_12_4096 sinft_12_4096 > gpurun_out/${R}_kernel_table_next.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 10 -c 10 -o gpurun_out/${R}_rlft3_full -f python tools/profile_rlft3.py 512 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 2 -c 2 -o gpurun_out/${R}_four1_20 -f python tools/profile_generic.py four1_20_64 > gpurun_out/ncu_four1_20.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_pass_kernel -s 1 -c 1 -o gpurun_out/${R}_four1_12 -f python tools/profile_generic.py four1_12_4096 > gpurun_out/ncu_four1_12.log 2>&1
cat gpurun_out/smoke.log; tail -n 2 gpurun_out/pytest_gpu.log
cat gpurun_out/${R}_kernel_table.txt | grep "^=="
