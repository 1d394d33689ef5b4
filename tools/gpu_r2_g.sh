#!/bin/bash
# 2-GPU call: push + pull exchange A/B, TMA default build checks, GPU test suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B=tools/cabi_bench
L=numrs_b200/libnumrs_b200.so
run() { echo "== $*"; timeout 60 $B $L "$@" 2>&1 | grep -v "^option" | grep -v "L512 \|L1024 \|pad_resp\|L4096 "; }
{
run rlft3:512
run rlft3:512 prefetch_dist=444
run rlft3:512 prefetch_dist=888
run rlft3:512 tma_col_mask=0
run fourn:512x512x512
run convlv:22:64
run convlv:22:64 tma_col_mask=0
} > gpurun_out/r02_g_tma_ab.txt 2>&1
cat gpurun_out/r02_g_tma_ab.txt
timeout 300 $TR --master-port 29551 tools/slab_modes.py 512 2.07 rlft3 push fused fused:1:1:2 fused:1:1:3 fused:1:1:5 fused:1:1:6 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee gpurun_out/r02_g_slab_modes_${N}gpu.txt
timeout 300 $TR --master-port 29552 tools/slab_modes.py 512 4.1 fourn push fused 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee -a gpurun_out/r02_g_slab_modes_${N}gpu.txt
timeout 300 $TR --master-port 29553 tools/slab_modes.py 1024 0 rlft3 push fused dma:4:1 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee -a gpurun_out/r02_g_slab_modes_${N}gpu.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_g_tests_${N}gpu.txt
