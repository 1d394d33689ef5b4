#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_j_tests_${N}gpu.txt
timeout 300 $TR --master-port 29581 tools/slab_modes.py 512 2.07 rlft3 push fused 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee gpurun_out/r02_j_slab_modes_${N}gpu.txt
