#!/bin/bash
# CPU-side memory / undefined-behaviour check of the DEVICE source and the planner: builds the test-only emulator
# (tests/emu) with AddressSanitizer or UBSan and runs the emulator test files on it.  Every "global memory" buffer of
# the emulated kernels is a malloc block and shared memory is a std::vector, so an out-of-bounds index in a kernel is
# a heap-buffer-overflow report.  Usage: tools/emu_sanitize.sh address|undefined
set -e
cd "$(dirname "$0")/.."
KIND=${1:-address}
D=$(mktemp -d)
FLAGS="-O1 -g -std=c++17 -fPIC -fopenmp -DNRB_EMU -fsanitize=$KIND -fno-sanitize-recover=undefined -Wno-unknown-pragmas"
g++ $FLAGS -c -o $D/emu_backend.o tests/emu/emu_backend.cpp &
g++ $FLAGS -x c++ -c -o $D/plan.o numrs_b200/csrc/plan.cpp &
g++ $FLAGS -x c++ -c -o $D/api.o numrs_b200/csrc/api.cpp &
wait
g++ -shared -fopenmp -fsanitize=$KIND -o $D/libnrb_emu.so $D/emu_backend.o $D/plan.o $D/api.o
make -C tests/emu -s
cp tests/emu/libnrb_emu.so $D/orig.so
trap 'cp $D/orig.so tests/emu/libnrb_emu.so; touch tests/emu/libnrb_emu.so; rm -rf $D' EXIT
cp $D/libnrb_emu.so tests/emu/libnrb_emu.so
touch tests/emu/libnrb_emu.so
PRE=""
[ "$KIND" = address ] && PRE=$(gcc -print-file-name=libasan.so)
LD_PRELOAD=$PRE ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  python -m pytest tests/test_emu_parity.py tests/test_emu_property.py tests/test_slab.py -x -q
