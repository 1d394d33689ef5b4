#!/bin/bash
# N-GPU experiment: pipelined (z-chunked) fused exchange, CTA cap of the peer-store pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
for cfg in "1 0" "2 148" "4 148" "2 0"; do
  set -- $cfg; c=$1; cap=$2
  NRB_XCHG_GRID_CAP=$cap timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2973$c bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --chunks $c > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_tmp.json').read()); print('  n=%d chunks=$c cap=$cap value %.0f GB/s  ms/step %.3f  roundtrip err %.2e' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roundtrip_rel_l2']))" || tail -5 gpurun_out/bench_tmp.err
done
