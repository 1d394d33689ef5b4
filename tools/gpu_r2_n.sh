#!/bin/bash
# compute-sanitizer on one small instance of every kernel family (round-2 kernels included)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py ) > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/r02_sanitizer_memcheck.txt
tail -6 gpurun_out/r02_sanitizer_memcheck.txt
( time timeout 700 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py ) > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck exit $?" >> gpurun_out/r02_sanitizer_racecheck.txt
tail -6 gpurun_out/r02_sanitizer_racecheck.txt
