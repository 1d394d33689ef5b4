#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
for cfg in "2 0" "2 148" "4 148"; do
  set -- $cfg; c=$1; cap=$2
  NRB_XCHG_GRID_CAP=$cap timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2972$c tools/slab_timeline.py 512 $c 2>&1 | grep -E "==|isign|Error|error" | head
done
