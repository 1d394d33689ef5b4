#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 4 8; do
    timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n tools/slab_breakdown.py 512 fused 2>&1 | grep -E "==|   |Error|error" | head -20
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 296$n$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench exit $?"; python -c "import sys,json; d=json.loads(open('gpurun_out/bench_n$n.json').read()); print('bench n=%d value %.0f GB/s  ms/step %.3f  e2e %.1f GB/s err %.2e' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roundtrip_rel_l2']))"
done
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29777 tools/slab_breakdown.py 1024 fused 2>&1 | grep -E "==|   |Error|error" | head -20
