#!/bin/bash
# Round 2, GPU batch 9: whole suite after the graph-capture / whole-step-K1 / GEMM changes, full bench line, fused GEMM on/off,
# compute-sanitizer over the hand-written kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_attack_success.py 2>&1 | tail -25 > $O/b9_pytest.log
tail -12 $O/b9_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/b9_bench.json 2> $O/b9_bench.err
tail -3 $O/b9_bench.err; python tools/bench_digest.py $O/b9_bench.json
for fg in 0 1; do
  DORPATCH_FUSED_GEMM=$fg timeout 400 python bench.py --precision bf16 --config c2 --steps 5 --warmup 3 --no-cpu-baseline --no-legs > $O/b9_c2_bf16_fg$fg.json 2> $O/b9_c2_bf16_fg$fg.err
  echo "== fused_gemm=$fg"; python tools/bench_digest.py $O/b9_c2_bf16_fg$fg.json | head -20
done
CS=/usr/local/cuda/bin/compute-sanitizer
for cfg in "memcheck bf16" "racecheck bf16" "memcheck fp32"; do
  set -- $cfg
  timeout 420 $CS --tool $1 --error-exitcode 9 --print-limit 10 python tools/sanitize_run.py $2 112 > $O/sanitize_$1_$2.log 2>&1
  echo "$1 $2 rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run ok' $O/sanitize_$1_$2.log | tr '\n' ' ')"
done
du -sh $O
