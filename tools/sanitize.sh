#!/bin/bash
# compute-sanitizer over the hand-written kernels (hygiene): memcheck, racecheck (shared-memory hazards in the
# mbarrier / cluster-barrier protocols), synccheck.  One GPU, a few minutes.  Output: gpurun_out/sanitize_*.log
#   gpurun --timeout 1200 -- 'bash tools/sanitize.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for prec in bf16 fp32; do
  for tool in memcheck racecheck synccheck; do
    timeout 900 $CS --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_run.py $prec 112 > $O/sanitize_${tool}_$prec.log 2>&1
    echo "$tool $prec rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run ok' $O/sanitize_${tool}_$prec.log | tr '\n' ' ')"
  done
done
