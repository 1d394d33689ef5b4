#!/bin/bash
# pipelined batch host calls: parity tests that go through them, then e2e of the batch workloads with and without
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch or full_size or convlv or correl or realft" 2>&1 | tail -3 | tee gpurun_out/r02_m_tests.txt
for w in four1_batch four1_1m convlv correl; do
  for p in 1 0; do
    NRB_PIPELINE_BATCHES=$p timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --workload $w > gpurun_out/r02_m_${w}_pipe$p.json 2> gpurun_out/r02_m_${w}_pipe$p.err
    python - gpurun_out/r02_m_${w}_pipe$p.json $w $p <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d["e2e"]
    print(sys.argv[2], "pipeline", sys.argv[3], "e2e", round(e["value"],1), "GB/s", round(e["ms_per_step"],2), "ms", "value", round(d["value"]))
except Exception as ex: print("unparsed", ex)
PY
  done
done 2>&1 | tee gpurun_out/r02_m_pipeline_ab.txt
