#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29591 tools/slab_modes.py 512 2.07 rlft3 fused fused:1:1:4:2 fused fused:1:1:4:2 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee gpurun_out/r02_k_slab_modes_${N}gpu.txt
timeout 300 $TR --master-port 29592 tools/slab_modes.py 512 4.1 fourn fused fused:1:1:4:2 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee -a gpurun_out/r02_k_slab_modes_${N}gpu.txt
timeout 300 $TR --master-port 29593 tools/slab_modes.py 1024 0 rlft3 fused fused:1:1:4:2 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee -a gpurun_out/r02_k_slab_modes_${N}gpu.txt
