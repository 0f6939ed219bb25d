#!/bin/bash
# Round 2, GPU batch 10: K1 launch-shape / store-path sweep (both precisions), op tests for the new statistics kernel and the
# accuracy-based fused-GEMM test, the redesigned attack-success tests, racecheck (light workload).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
for prec in bf16 tf32; do
  timeout 400 python tools/k1_step_sweep.py $prec $O/b10_k1_sweep_$prec.jsonl 2>&1 | grep -E "BEST|auto|Error|error" | head -30
done
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fused_gemm.py -m gpu -q -s 2>&1 | grep -E "fused|logits|cosine|passed|failed|Error" | head -20
#timeout 900 python -m pytest tests/test_gpu_attack_success.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|skipped|assert" > $O/b10_attack_success.log
#cat $O/b10_attack_success.log | head -90
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 500 $CS --tool racecheck --error-exitcode 9 --print-limit 10 python tools/sanitize_run.py bf16 112 light > $O/sanitize_racecheck_bf16.log 2>&1
echo "racecheck bf16 rc=$? : $(grep -E 'RACECHECK SUMMARY|sanitize_run ok|Error' $O/sanitize_racecheck_bf16.log | tr '\n' ' ')"
timeout 300 $CS --tool synccheck --error-exitcode 9 --print-limit 10 python tools/sanitize_run.py bf16 112 light > $O/sanitize_synccheck_bf16.log 2>&1
echo "synccheck bf16 rc=$? : $(grep -E 'ERROR SUMMARY|sanitize_run ok|Error' $O/sanitize_synccheck_bf16.log | tr '\n' ' ')"
du -sh $O
