#!/bin/bash
# Round 2, GPU batch 11: per-layer time of the tcgen05 GEMM (gemmbench), K1 sweep with the new launch rule
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o $O/gemmbench tools/gemmbench.cu -Ldorpatch_b200/lib -ldorpatch -Xlinker -rpath,$PWD/dorpatch_b200/lib 2> $O/gemmbench_build.log || cat $O/gemmbench_build.log
timeout 300 $O/gemmbench 256 > $O/b11_gemmbench.log 2>&1; cat $O/b11_gemmbench.log; rm -f $O/gemmbench
for prec in bf16 tf32; do
  timeout 400 python tools/k1_step_sweep.py $prec $O/b11_k1_sweep_$prec.jsonl 2>&1 | grep -E "BEST|auto|Error|error" | head -30
done
