#!/bin/bash
# Round-2 evidence run on ONE GPU: smoke, the whole GPU test tier, both bench arms, every workload's bench line, per-kernel
# tables, the ncu launch list of the bench command.  Usage: gpurun --timeout 2400 -- ./tools/gpu_r2_final.sh [tag]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r02_final}
B=tools/cabi_bench
L=numrs_b200/libnumrs_b200.so
run() { echo "== $*"; timeout 60 $B $L "$@" 2>&1 | grep -v "^option" | grep -v "L512 \|L1024 \|pad_resp\|L4096 "; }
{ run convlv:22:64; run convlv:22:64 mid_prefetch=148; run convlv:22:64 mid_prefetch=296; run correl:22:64 mid_prefetch=148; } > gpurun_out/${R}_mid_prefetch_ab.txt 2>&1
grep "one call\|^==" gpurun_out/${R}_mid_prefetch_ab.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${R}_smoke.log; tail -2 gpurun_out/${R}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${R}_pytest_gpu.log; tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench_rlft3_512.json 2> gpurun_out/${R}_bench_ours.err; echo "bench exit $?"
for w in fourn3d_512 four1_batch four1_1m fourn2d convlv correl; do
  timeout 400 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/${R}_bench_$w.json 2> gpurun_out/${R}_bench_$w.err; echo "$w exit $?"
done
timeout 300 python tools/kernel_table.py rlft3_512 four1_12_4096 four1_20_64 fourn2d_8192 convlv_22_16 correl_22_16 > gpurun_out/${R}_kernel_table.txt 2>&1
timeout 300 python tools/kernel_table.py twofft_20_16 twofft_12_4096 correlnorm_22_16 correlnormfast_22_16 autocorrel_22_16 cosft1_22_16 cosft1_12_4096 cosft2_22_16 cosft2_12_4096 sinft_12_4096 > gpurun_out/${R}_kernel_table_next.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${R}_ncu_bench.log 2>&1
for f in gpurun_out/${R}_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","roundtrip_rel_l2","gpu_launches")}, (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
except Exception as e: print("unparsed", e)
PY
done
