#!/bin/bash
# Round 2, GPU batch 7: details of the failing tests of batch 6 (attack-success bits, graph capture, fused GEMM engine test),
# GroupNorm v2 per-shape table (gnbench), fused tcgen05 GEMM on/off at bf16 c2, small ncu counters of the GN kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_attack_success.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -60 > $O/b7_attack_success.log
cat $O/b7_attack_success.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k graph 2>&1 | tail -15
DORPATCH_TEST_FUSED_GEMM=1 timeout 300 python -m pytest tests/test_gpu_fused_gemm.py -m gpu -q 2>&1 | grep -E "assert|Error|passed|failed" | head -20
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o $O/gnbench tools/gnbench.cu -Ldorpatch_b200/lib -ldorpatch -Xlinker -rpath,$PWD/dorpatch_b200/lib 2> $O/gnbench_build.log || cat $O/gnbench_build.log
DORPATCH_GN=v2 timeout 600 $O/gnbench 256 > $O/b7_gnbench_v2.log 2>&1; cat $O/b7_gnbench_v2.log
rm -f $O/gnbench
for fg in 0 1; do
  DORPATCH_FUSED_GEMM=$fg timeout 400 python bench.py --precision bf16 --config c2 --steps 5 --warmup 3 --no-cpu-baseline --no-legs > $O/b7_c2_bf16_fg$fg.json 2> $O/b7_c2_bf16_fg$fg.err
  echo "== fused_gemm=$fg"; python tools/bench_digest.py $O/b7_c2_bf16_fg$fg.json | head -22
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size,launch__cluster_size \
  --clock-control none --profile-from-start off -k regex:"gn2::|expand_kernel|stem_" --csv --log-file $O/b7_gn_counters_bf16_c2.csv \
  python bench.py --ncu --warmup 3 --precision bf16 --config c2 > $O/b7_ncu_c.log 2>&1
tail -2 $O/b7_ncu_c.log
du -sh $O
