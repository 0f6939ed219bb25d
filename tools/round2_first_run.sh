#!/bin/bash
# First GPU batch of round 2 (one gpurun call, ~6 GPU-minutes): everything queued at the end of round 1 that
# needs hardware.  Outputs land in gpurun_out/ (scratch); copy what should be judged into profiles/.
#   gpurun --timeout 900 -- 'bash tools/round2_first_run.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out

# 0. access-pattern micro-benchmarks: where does the ~3.1 TB/s ceiling of the GroupNorm kernels come from?
timeout 60 nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o $O/membench tools/membench.cu > $O/r2_membench.log 2>&1 \
  && timeout 120 $O/membench >> $O/r2_membench.log 2>&1
cat $O/r2_membench.log; rm -f $O/membench

# 1. first hardware run of the tcgen05 GroupNorm-prologue GEMM (kernels_gemm.cu); a hang must not take the box down
DORPATCH_TEST_FUSED_GEMM=1 timeout 180 python -m pytest tests/test_gpu_fused_gemm.py -m gpu -x -q > $O/r2_fused_gemm_test.log 2>&1
tail -5 $O/r2_fused_gemm_test.log

# 2. if it passes: what it buys (same bench, fused path on)
if grep -q " passed" $O/r2_fused_gemm_test.log && ! grep -q "failed" $O/r2_fused_gemm_test.log; then
  DORPATCH_FUSED_GEMM=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_bench_fused_gemm.json 2> $O/r2_bench_fused_gemm.err
  python - <<'PY'
import json
a = json.load(open("gpurun_out/r2_bench_fused_gemm.json"))
print("fused gemm:", round(a["value"]), "samples/s", {k: v["ms"] for k, v in list(a.get("kernels", {}).items())[:8]})
PY
fi

# 3. source-level profile of the top kernel (GroupNorm backward), the way K1's avoidable instructions were found:
#    ncu -i gpurun_out/r2_gn_bwd.ncu-rep --page source --csv   (SASS with per-instruction samples)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_bwd_cluster --launch-skip 4 -c 3 \
  -o $O/r2_gn_bwd python bench.py --ncu --warmup 3 > $O/r2_ncu_gn_bwd.log 2>&1
tail -2 $O/r2_ncu_gn_bwd.log

# 4. GroupNorm launch-shape variants on the final code (one bench line each)
for cfg in "DORPATCH_GN_CL16=1" "DORPATCH_GN_BIGTHREADS=256" "DORPATCH_GN=twopass"; do
  env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_bench_${cfg%%=*}.json 2>> $O/r2_bench_variants.err
  python - "$cfg" <<'PY'
import json, sys
cfg = sys.argv[1]
a = json.load(open("gpurun_out/r2_bench_%s.json" % cfg.split("=")[0]))
k = a.get("kernels", {})
print(cfg, round(a["value"]), {n: k[n]["ms"] for n in ("gn_relu_bwd", "gn_relu_fwd") if n in k})
PY
done

# 5. small chunks with one lane: does an L2-resident working set pay?
for ch in 16 32; do
  DORPATCH_LANES=1 timeout 300 python bench.py --steps 3 --warmup 3 --chunk $ch --no-cpu-baseline > $O/r2_bench_chunk$ch.json 2>> $O/r2_bench_variants.err
  python -c "
import json; a = json.load(open('gpurun_out/r2_bench_chunk$ch.json')); print('chunk $ch lanes 1:', round(a['value']), a['gpu_launches'])"
done
