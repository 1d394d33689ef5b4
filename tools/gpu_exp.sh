#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 8 4; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r01_bench_rlft3_512_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench exit $?"; python -c "import sys,json; d=json.loads(open('gpurun_out/r01_bench_rlft3_512_n$n.json').read()); print('bench n=%d value %.0f GB/s  ms/step %.3f  e2e %.1f GB/s err %.2e clocks %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roundtrip_rel_l2'], d['clocks']))"
done
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 tools/slab_breakdown.py 512 fused flags 2>&1 | grep -E "==|   |Error|error" | head -20
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/slab_breakdown.py 1024 fused flags 2>&1 | grep -E "==|   |Error|error" | head -20
