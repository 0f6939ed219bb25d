#!/bin/bash
# one-off measurement batch: K1 packed-mask rewrite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/pytest_b.log 2>&1
tail -3 gpurun_out/pytest_b.log
: > gpurun_out/k1_sweep_b.jsonl
for cfg in "X=1" "DORPATCH_K1_CTAS=6" "DORPATCH_K1_CTAS=4" "DORPATCH_K1_ROWS=4 DORPATCH_K1_CTAS=6" "DORPATCH_K1_ROWS=16" "K1_PRECISION=tf32"; do
  env $cfg timeout 120 python tools/k1_sweep.py >> gpurun_out/k1_sweep_b.jsonl 2>> gpurun_out/k1_sweep_b.err
done
cat gpurun_out/k1_sweep_b.jsonl
