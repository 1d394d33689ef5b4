#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for b in flags nccl; do
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2951$NG tools/slab_breakdown.py 512 fused $b 2>&1 | grep -E "==|   |Error|error" | head -20
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2961$NG bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n$NG.err; python -c "import sys,json; d=json.loads(open('gpurun_out/bench_n$NG.json').read()); print('bench n=%d value %.0f GB/s  ms/step %.3f  e2e %.1f GB/s err %.2e' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roundtrip_rel_l2']))"
