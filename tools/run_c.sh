#!/bin/bash
# round-1 closing batch: full GPU suite, bench line, K1 ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_c.log 2>&1
tail -3 gpurun_out/pytest_c.log
timeout 400 python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
tail -c 300 gpurun_out/bench_c.err; head -c 400 gpurun_out/bench_c.json; echo
timeout 200 ncu --set full --clock-control none --import-source on -k regex:expand_kernel --launch-skip 10 -c 1 -o gpurun_out/prof_k1_c python tools/k1_sweep.py > gpurun_out/ncu_k1_c.log 2>&1
tail -2 gpurun_out/ncu_k1_c.log
