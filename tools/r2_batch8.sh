#!/bin/bash
# Round 2, GPU batch 8: L2-windowed two-kernel ceiling (tools/l2bench.cu); attack-success rates per precision (full output)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o $O/l2bench tools/l2bench.cu && timeout 120 $O/l2bench > $O/b8_l2bench.log 2>&1
rm -f $O/l2bench; cat $O/b8_l2bench.log
timeout 600 python -m pytest tests/test_gpu_attack_success.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed" > $O/b8_attack_success.log
cat $O/b8_attack_success.log
