#!/bin/bash
# bench.py lines at N = 2 and N = 4 on a 4-GPU box (the driver's SCALE run does the same at round end)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for N in 2 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N"
  timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_scale_bench_n$N.json 2> gpurun_out/r02_scale_bench_n$N.err; tail -c 200 gpurun_out/r02_scale_bench_n$N.err
  timeout 300 $TR bench.py --gpus $N --workload fourn3d_512 --steps 10 --warmup 3 > gpurun_out/r02_scale_fourn3d_n$N.json 2> gpurun_out/r02_scale_fourn3d_n$N.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29579 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r02_scale_reference_n4.json 2> gpurun_out/r02_scale_reference_n4.err
for f in gpurun_out/r02_scale_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","roundtrip_rel_l2","gpu_launches")}, (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"))
except Exception as e: print("unparsed", e)
PY
done
