#!/bin/bash
# Round 2, GPU batch 2: GroupNorm v2 kernels -- correctness + bandwidth per layer shape against v1, then the GPU kernel tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o $O/gnbench tools/gnbench.cu -Ldorpatch_b200/lib -ldorpatch -Xlinker -rpath,$PWD/dorpatch_b200/lib 2> $O/gnbench_build.log || { cat $O/gnbench_build.log; exit 1; }
for v in v2 v1; do
  DORPATCH_GN=$v timeout 600 $O/gnbench 256 > $O/r2_gnbench_$v.log 2>&1
  echo "== $v"; cat $O/r2_gnbench_$v.log
done
for var in "DORPATCH_GN2_SOFT=104" "DORPATCH_GN2_MAXCL=8" "DORPATCH_GN2_DYS_BIG=0"; do
  for c in 64 256; do
    echo "== v2 $var C=$c"
    env $var DORPATCH_GN=v2 timeout 300 $O/gnbench 256 $c 2>&1 | tee -a $O/r2_gnbench_variants.log
  done
done
rm -f $O/gnbench
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py -m gpu -x -q > $O/r2_b2_pytest.log 2>&1
tail -5 $O/r2_b2_pytest.log
