#!/bin/bash
# Round 2, GPU batch 5: GN v2 with the shuffle reduce (trace + full table), graph-capture diagnosis, K1 in-step launch-shape sweep,
# tcgen05 fused GEMM on/off at bf16 c2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o $O/gnbench tools/gnbench.cu -Ldorpatch_b200/lib -ldorpatch -Xlinker -rpath,$PWD/dorpatch_b200/lib 2> $O/gnbench_build.log || { cat $O/gnbench_build.log; exit 1; }
for c in 64 256; do GNBENCH_TRACE=1 DORPATCH_GN=v2 timeout 300 $O/gnbench 256 $c 2>&1 | grep -v NEG | tee -a $O/r2_gnbench_trace2.log; done
DORPATCH_GN=v2 timeout 600 $O/gnbench 256 > $O/r2_gnbench_v2b.log 2>&1; cat $O/r2_gnbench_v2b.log
rm -f $O/gnbench
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "graph" 2>&1 | tail -8
for prec in bf16 tf32; do
  for rows in auto 4 7 8 14 16; do for sg in auto 1 2; do
    if [ $rows = auto ] && [ $sg != auto ]; then continue; fi
    if [ $rows != auto ] && [ $sg = auto ]; then continue; fi
    if [ $rows = auto ]; then timeout 120 python tools/k1_step_sweep.py $prec 256 16 2>/dev/null | tail -1
    else DORPATCH_K1_ROWS=$rows DORPATCH_K1_SG=$sg timeout 120 python tools/k1_step_sweep.py $prec 256 16 2>/dev/null | tail -1; fi
  done; done
done | tee $O/r2_k1_step_sweep.jsonl
for fg in 0 1; do
  DORPATCH_FUSED_GEMM=$fg timeout 400 python bench.py --precision bf16 --config c2 --steps 5 --warmup 3 --no-cpu-baseline --no-legs > $O/r2_b5_c2_bf16_fg$fg.json 2> $O/r2_b5_c2_bf16_fg$fg.err
  python - $fg <<'PY'
import json, sys
try:
    a = json.load(open("gpurun_out/r2_b5_c2_bf16_fg%s.json" % sys.argv[1]))
    print("fused_gemm", sys.argv[1], "value", round(a["value"]), "ms", round(a["ms_per_step"], 2), {n: v["ms"] for n, v in list(a["kernels"].items())[:10]}, "K1", round(a["roofline"]["frac"], 3))
except Exception as ex:
    print("bench parse failed", ex)
PY
done
