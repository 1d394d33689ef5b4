#!/bin/bash
# 2-GPU experiment: NVLink store patterns + the new rows' GPU parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 120 ./tools/p2p_store_bench > gpurun_out/p2p_store_bench.log 2>&1; echo "exit $?" >> gpurun_out/p2p_store_bench.log
cat gpurun_out/p2p_store_bench.log
timeout 900 python -m pytest tests -m gpu -x -q -k "twofft or correl_normalized or autocorrel_fast or spectrum or host_mirror" > gpurun_out/pytest_next.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_next.log
tail -5 gpurun_out/pytest_next.log
