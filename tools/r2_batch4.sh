#!/bin/bash
# Round 2, GPU batch 4: GroupNorm v3 (streaming two-phase, second read from L2) against v2; op-level tests incl. the tcgen05 GEMM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o $O/gnbench tools/gnbench.cu -Ldorpatch_b200/lib -ldorpatch -Xlinker -rpath,$PWD/dorpatch_b200/lib 2> $O/gnbench_build.log || { cat $O/gnbench_build.log; exit 1; }
DORPATCH_GN=v3 timeout 600 $O/gnbench 256 > $O/r2_gnbench_v3.log 2>&1
echo "== v3"; cat $O/r2_gnbench_v3.log
for var in "DORPATCH_GN3_ITER=8" "DORPATCH_GN3_L2MB=8" "DORPATCH_GN3_L2MB=64"; do
  for c in 64 256 1024; do
    echo "== v3 $var C=$c"
    env $var DORPATCH_GN=v3 timeout 300 $O/gnbench 256 $c 2>&1 | grep -v NEG | tee -a $O/r2_gnbench_v3_variants.log
  done
done
rm -f $O/gnbench
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -15 > $O/r2_b4_pytest.log
cat $O/r2_b4_pytest.log
