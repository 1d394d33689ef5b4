#!/bin/bash
# 8-GPU call: push + pull exchange at 8 ranks (parity + timing), bench lines for the scaling table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
ls /sys/devices/system/node | grep -c "^node"
timeout 300 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -k "fused_exchange and 8-64x128x256" 2>&1 | tail -3 | tee gpurun_out/r02_h_tests_${N}gpu.txt
timeout 300 $TR --master-port 29561 tools/slab_modes.py 512 2.09 rlft3 push fused fused:1:1:3 fused:1:1:5 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee gpurun_out/r02_h_slab_modes_${N}gpu.txt
timeout 300 $TR --master-port 29562 tools/slab_modes.py 512 4.2 fourn push fused 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee -a gpurun_out/r02_h_slab_modes_${N}gpu.txt
timeout 300 $TR --master-port 29563 tools/slab_modes.py 1024 0 rlft3 push fused fused:1:1:3 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tee -a gpurun_out/r02_h_slab_modes_${N}gpu.txt
timeout 300 $TR --master-port 29564 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_h_bench_n$N.json 2> gpurun_out/r02_h_bench_n$N.err; tail -c 300 gpurun_out/r02_h_bench_n$N.err
timeout 400 $TR --master-port 29565 bench.py --gpus $N --workload rlft3_1024 --steps 6 --warmup 3 > gpurun_out/r02_h_rlft3_1024_n$N.json 2> gpurun_out/r02_h_rlft3_1024_n$N.err; tail -c 300 gpurun_out/r02_h_rlft3_1024_n$N.err
timeout 300 $TR --master-port 29566 bench.py --gpus $N --workload fourn3d_512 --steps 10 --warmup 3 > gpurun_out/r02_h_fourn3d_n$N.json 2> gpurun_out/r02_h_fourn3d_n$N.err; tail -c 300 gpurun_out/r02_h_fourn3d_n$N.err
timeout 300 $TR --master-port 29567 bench.py --gpus $N --workload four1_batch --steps 10 --warmup 3 > gpurun_out/r02_h_four1b_n$N.json 2> gpurun_out/r02_h_four1b_n$N.err; tail -c 300 gpurun_out/r02_h_four1b_n$N.err
timeout 400 $TR --master-port 29568 bench.py --gpus $N --workload convlv --steps 5 --warmup 3 > gpurun_out/r02_h_convlv_n$N.json 2> gpurun_out/r02_h_convlv_n$N.err; tail -c 300 gpurun_out/r02_h_convlv_n$N.err
for f in gpurun_out/r02_h_*n$N.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","roundtrip_rel_l2","gpu_launches")}, d.get("e2e"))
except Exception as e: print("unparsed", e)
PY
done
