#!/bin/bash
# Round-2 multi-GPU check (run with gpurun --gpus 2): the new GPU tests, then bench lines at N = 1 and N = 2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader; nproc; free -g | sed -n 2p
nvidia-smi topo -m 2>/dev/null | head -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
{
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_host_cpp.py -x -q -m gpu 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -5
} 2>&1 | tee gpurun_out/r02_a_tests.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_a_bench_n1.json 2> gpurun_out/r02_a_bench_n1.err; tail -c 600 gpurun_out/r02_a_bench_n1.err
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_a_bench_n$N.json 2> gpurun_out/r02_a_bench_n$N.err; tail -c 600 gpurun_out/r02_a_bench_n$N.err
timeout 300 python bench.py --workload fourn3d_512 --steps 10 --warmup 3 > gpurun_out/r02_a_fourn3d_n1.json 2> gpurun_out/r02_a_fourn3d_n1.err; tail -c 600 gpurun_out/r02_a_fourn3d_n1.err
timeout 300 $TR bench.py --gpus $N --workload fourn3d_512 --steps 10 --warmup 3 > gpurun_out/r02_a_fourn3d_n$N.json 2> gpurun_out/r02_a_fourn3d_n$N.err; tail -c 600 gpurun_out/r02_a_fourn3d_n$N.err
timeout 300 $TR bench.py --gpus $N --workload four1_batch --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_a_four1b_n$N.json 2> gpurun_out/r02_a_four1b_n$N.err; tail -c 600 gpurun_out/r02_a_four1b_n$N.err
for f in gpurun_out/r02_a_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","roundtrip_rel_l2","gpu_launches")}, d.get("e2e"), (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print("unparsed", e)
PY
done
