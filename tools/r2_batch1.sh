#!/bin/bash
# Round 2, GPU batch 1: first hardware run of the tcgen05 GEMM, c3 (64 x 32) baselines at tf32 / bf16 with the
# per-category breakdown, and a source-level ncu capture of the GroupNorm kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader

# 1. tcgen05 GN-prologue GEMM: a hang must not take the box down
DORPATCH_TEST_FUSED_GEMM=1 timeout 240 python -m pytest tests/test_gpu_fused_gemm.py -m gpu -x -q > $O/r2_fused_gemm_test.log 2>&1
echo "fused gemm test rc=$?"; tail -15 $O/r2_fused_gemm_test.log

# 2. c3 baselines (64 images x 32 EOT)
for prec in tf32 bf16; do
  timeout 400 python bench.py --precision $prec --batch 64 --eot 32 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_b1_c3_$prec.json 2> $O/r2_b1_c3_$prec.err
  python - $prec <<'PY'
import json, sys
p = sys.argv[1]
try:
    a = json.load(open("gpurun_out/r2_b1_c3_%s.json" % p))
    k = a.get("kernels", {})
    print(p, "c3:", round(a["value"]), "samples/s", round(a["ms_per_step"], 1), "ms; e2e", round(a["e2e"]["value"]),
          {n: (v["ms"], v.get("frac_of_hbm_peak", v.get("frac_of_bf16_peak"))) for n, v in list(k.items())[:10]})
except Exception as ex:
    print(p, "failed", ex)
PY
done
tail -3 $O/r2_b1_c3_tf32.err

# 3. ncu of the GroupNorm kernels (bf16), source-level
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gn_bwd_cluster --launch-skip 40 -c 6 \
  -o $O/r2_gn_bwd -f python bench.py --ncu --warmup 3 > $O/r2_ncu_gn_bwd.log 2>&1
tail -2 $O/r2_ncu_gn_bwd.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gn_fwd_cluster --launch-skip 40 -c 6 \
  -o $O/r2_gn_fwd -f python bench.py --ncu --warmup 3 > $O/r2_ncu_gn_fwd.log 2>&1
tail -2 $O/r2_ncu_gn_fwd.log
ls -la $O/*.ncu-rep
