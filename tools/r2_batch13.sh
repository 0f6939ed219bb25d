#!/bin/bash
# Round 2, GPU batch 13: end-metric parity tests (frozen images + noise floor [+ saturated when its fixture exists]), whole suite, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_attack_success.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|skipped|assert|Error" > $O/b13_attack_success.log
cat $O/b13_attack_success.log | head -120
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_attack_success.py 2>&1 | tail -8 > $O/b13_pytest.log
cat $O/b13_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/b13_bench.json 2> $O/b13_bench.err
tail -3 $O/b13_bench.err; python tools/bench_digest.py $O/b13_bench.json 2>/dev/null | head -60
