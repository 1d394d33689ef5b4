#!/bin/bash
# 8-GPU call: multi-rank parity at 8 ranks, exchange-mode comparison at 512^3 and 1024^3, the official bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc; free -g | sed -n 2p
{
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -k "8- or all_devices or nccl" 2>&1 | tail -6
} 2>&1 | tee gpurun_out/r02_d_tests_${N}gpu.txt
timeout 300 $TR --master-port 29541 tools/slab_modes.py 512 2.14 rlft3 fused fused:2 fused:4 dma:1:1 dma:1:4 dma:2:1 dma:2:4 dma:4:4 nccl 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee gpurun_out/r02_d_slab_modes_${N}gpu.txt
timeout 300 $TR --master-port 29542 tools/slab_modes.py 1024 0 rlft3 fused dma:1:4 dma:2:4 dma:4:4 dma:4:1 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee -a gpurun_out/r02_d_slab_modes_${N}gpu.txt
timeout 300 $TR --master-port 29543 tools/slab_modes.py 512 4.3 fourn fused dma:2:4 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee -a gpurun_out/r02_d_slab_modes_${N}gpu.txt
timeout 300 $TR --master-port 29544 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_d_bench_n$N.json 2> gpurun_out/r02_d_bench_n$N.err; tail -c 400 gpurun_out/r02_d_bench_n$N.err
timeout 400 $TR --master-port 29545 bench.py --gpus $N --workload rlft3_1024 --steps 6 --warmup 3 > gpurun_out/r02_d_rlft3_1024_n$N.json 2> gpurun_out/r02_d_rlft3_1024_n$N.err; tail -c 400 gpurun_out/r02_d_rlft3_1024_n$N.err
for f in gpurun_out/r02_d_*n$N.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","roundtrip_rel_l2","gpu_launches")}, d.get("e2e"))
except Exception as e: print("unparsed", e)
PY
done
