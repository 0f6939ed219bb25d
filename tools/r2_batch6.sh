#!/bin/bash
# Round 2, GPU batch 6 (re-entry baseline): whole GPU suite, tcgen05 GEMM opt-in test, full bench line, launch lists of the headline
# step (tf32 c3) and the bf16 c2 step, ncu --set full of the GroupNorm v2 + K1 kernels inside a bf16 c2 step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > $O/b6_gpu.txt; nproc >> $O/b6_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/b6_pytest.log
tail -12 $O/b6_pytest.log
DORPATCH_TEST_FUSED_GEMM=1 timeout 300 python -m pytest tests/test_gpu_fused_gemm.py -m gpu -q > $O/b6_fused_gemm_test.log 2>&1
tail -6 $O/b6_fused_gemm_test.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/b6_bench.json 2> $O/b6_bench.err
tail -3 $O/b6_bench.err
python tools/bench_digest.py $O/b6_bench.json
# launch lists (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/b6_launches_tf32_c3.csv \
  python bench.py --ncu --warmup 3 > $O/b6_ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/b6_launches_bf16_c2.csv \
  python bench.py --ncu --warmup 3 --precision bf16 --config c2 > $O/b6_ncu_l2.log 2>&1
python tools/launch_summary.py $O/b6_launches_tf32_c3.csv "tf32 c3 step" | head -40
python tools/launch_summary.py $O/b6_launches_bf16_c2.csv "bf16 c2 step" | head -40
# full counters of the hand-written hot kernels in a bf16 c2 step
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"fwd_kernel|bwd_kernel|expand_kernel|stem_" -c 48 \
  -o $O/b6_hot_bf16 python bench.py --ncu --warmup 3 --precision bf16 --config c2 > $O/b6_ncu_f1.log 2>&1
tail -2 $O/b6_ncu_f1.log
python tools/ncu_summary.py $O/b6_hot_bf16.ncu-rep "bf16 c2 step, hand-written hot kernels" > $O/b6_hot_bf16.txt 2>&1
grep -c duration $O/b6_hot_bf16.txt
