#!/bin/bash
# final 8-GPU check with the round's last code: one 8-rank parity test, the bench line, the reference arm under torchrun
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -k "fused_exchange and 8-64x128x256" 2>&1 | tail -2 | tee gpurun_out/r02_q_tests_${N}gpu.txt
timeout 300 $TR --master-port 29601 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_q_bench_n$N.json 2> gpurun_out/r02_q_bench_n$N.err; tail -c 300 gpurun_out/r02_q_bench_n$N.err
timeout 300 $TR --master-port 29602 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02_q_reference_n$N.json 2> gpurun_out/r02_q_reference_n$N.err
for f in gpurun_out/r02_q_*n$N.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","roundtrip_rel_l2","gpu_launches")}, (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"))
except Exception as e: print("unparsed", e)
PY
done
