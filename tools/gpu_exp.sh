#!/bin/bash
# sanitizer evidence for the kernels added after the first sanitizer run + C4 bench refresh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log
tail -4 gpurun_out/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck.log
tail -4 gpurun_out/sanitizer_racecheck.log
for w in convlv correl; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/r01_bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 300 python tools/kernel_table.py convlv_22_16 correl_22_16 autocorrel_22_16 correlnorm_22_16 correlnormfast_22_16 > gpurun_out/r01_kernel_table_c4.txt 2>&1
grep "^==" gpurun_out/r01_kernel_table_c4.txt; grep spectral gpurun_out/r01_kernel_table_c4.txt
