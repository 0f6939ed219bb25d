#!/bin/bash
# one-off measurement batch (K1 launch policy sweep, chunked GroupNorm A/B)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/pytest_a.log 2>&1
tail -3 gpurun_out/pytest_a.log
: > gpurun_out/k1_sweep.jsonl
for cfg in "" "DORPATCH_K1_BALANCE=0" "DORPATCH_K1_SG=1" "DORPATCH_K1_SG=2" "DORPATCH_K1_ROWS=4" "DORPATCH_K1_ROWS=16" "DORPATCH_K1_ROWS=4 DORPATCH_K1_SG=1" "DORPATCH_K1_ROWS=16 DORPATCH_K1_SG=2"; do
  env $cfg timeout 120 python tools/k1_sweep.py >> gpurun_out/k1_sweep.jsonl 2>> gpurun_out/k1_sweep.err
done
cat gpurun_out/k1_sweep.jsonl
for cfg in "X=1" "DORPATCH_GN_CHUNKED=0" "DORPATCH_GN_CL16=1"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a_${cfg%%=*}.json 2>> gpurun_out/bench_a.err
  python - <<PY
import json
a=json.load(open("gpurun_out/bench_a_${cfg%%=*}.json"))
k=a.get("kernels",{})
print(round(a["value"]), round(a["ms_per_step"],2), round(a["roofline"]["frac"],3), {n:v["ms"] for n,v in list(k.items())[:6]})
PY
done
