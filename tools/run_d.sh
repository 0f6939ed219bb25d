#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/pytest_d.log 2>&1
tail -15 gpurun_out/pytest_d.log
timeout 300 python bench.py > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
tail -c 300 gpurun_out/bench_d.err; python -c "
import json;a=json.load(open('gpurun_out/bench_d.json'));print(a['value'],a['roofline'])"
