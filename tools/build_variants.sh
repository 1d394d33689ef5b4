#!/bin/bash
# Builds the library variants the next A/B needs into variants/ (git-ignored, travels with gpurun):
#   lib_tw3.so     radix-8 twiddle powers from three table loads          (-DNRB_TW_LOADS=3)
#   lib_xpose.so   cheap addressing for the transposing 1024-point pass   (-DNRB_SIMPLE_XPOSE_MASK=1024)
#   lib_simple.so  cheap addressing for more line lengths: contiguous 2048 / 4096 / 8192, strided 256 / 512 / 1024
set -e
cd "$(dirname "$0")/../numrs_b200/csrc"
mkdir -p ../../variants
build() { name=$1; shift; make -s -j8 OBJDIR=build_$name OUT=../../variants/lib_$name.so EXTRA="$*"; echo "built variants/lib_$name.so ($*)"; }
build tw3 -DNRB_TW_LOADS=3
build xpose -DNRB_SIMPLE_XPOSE_MASK=1024
build simple "-DNRB_SIMPLE_ROW_MASK=0x3800 -DNRB_SIMPLE_COL_MASK=0x700"
