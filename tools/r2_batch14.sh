#!/bin/bash
# Round 2, GPU batch 14: GroupNorm v2 plan variants (cluster cap, slab budget) on every layer shape; ncu --set full of the GN v2 and
# K1 kernels in a bf16 c2 step (few launches each, small report)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o $O/gnbench tools/gnbench.cu -Ldorpatch_b200/lib -ldorpatch -Xlinker -rpath,$PWD/dorpatch_b200/lib 2> $O/gnbench_build.log || cat $O/gnbench_build.log
for var in "DORPATCH_GN2_MAXCL=16" "DORPATCH_GN2_MAXCL=8" "DORPATCH_GN2_MAXCL=4" "DORPATCH_GN2_MAXCL=2" "DORPATCH_GN2_SOFT=72" "DORPATCH_GN2_SOFT=224"; do
  echo "== $var"
  env $var GNBENCH_DTYPE=b timeout 200 $O/gnbench 256 2>&1 | grep -v NEG | awk '{print $2,$4,"fwd",$10,"bwd",$16,"bwd+add",$22}'
done > $O/b14_gn_variants.log 2>&1
cat $O/b14_gn_variants.log
rm -f $O/gnbench
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"^fwd_kernel|^bwd_kernel" --launch-skip 10 -c 10 \
  -o $O/b14_gn_bf16 python bench.py --ncu --warmup 3 --precision bf16 --config c2 > $O/b14_ncu1.log 2>&1
tail -2 $O/b14_ncu1.log
python tools/ncu_summary.py $O/b14_gn_bf16.ncu-rep "GroupNorm v2 kernels inside a bf16 c2 step (launches 11-20 of the step)" > $O/b14_gn_bf16.txt 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"expand_kernel|stem_|maxpool" -c 6 \
  -o $O/b14_k1_bf16 python bench.py --ncu --warmup 3 --precision bf16 --config c2 > $O/b14_ncu2.log 2>&1
python tools/ncu_summary.py $O/b14_k1_bf16.ncu-rep "K1 (whole-step launch), stem and max-pool kernels inside a bf16 c2 step" > $O/b14_k1_bf16.txt 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"expand_kernel|^fwd_kernel|^bwd_kernel" -c 8 \
  -o $O/b14_tf32 python bench.py --ncu --warmup 3 --precision tf32 --config c3 > $O/b14_ncu3.log 2>&1
python tools/ncu_summary.py $O/b14_tf32.ncu-rep "K1 (whole-step launch) and the first GroupNorm kernels inside a tf32 c3 step" > $O/b14_tf32.txt 2>&1
grep -c duration $O/b14_*.txt; du -sh $O; ls -la $O/*.ncu-rep
