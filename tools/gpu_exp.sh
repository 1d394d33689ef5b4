#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
for c in 1 4; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2975$c tools/slab_dma_timeline.py 512 $c 2>&1 | grep -E "==|Error|error" | head
done
for cfg in "dma 1" "dma 4"; do
  set -- $cfg; x=$1; c=$2
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2976$c bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --exchange $x --chunks $c > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_tmp.json').read()); print('  n=%d exchange=$x chunks=$c value %.0f GB/s  ms/step %.3f  roundtrip err %.2e' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roundtrip_rel_l2']))" || tail -5 gpurun_out/bench_tmp.err
done
