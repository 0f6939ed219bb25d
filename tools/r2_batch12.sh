#!/bin/bash
# Round 2, GPU batch 12: classifier chunk size with CUDA graphs on (does an L2-resident working set pay now that launches are free?)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
for ch in 32 64 128 256 512; do
  timeout 300 python bench.py --precision bf16 --config c2 --chunk $ch --steps 6 --warmup 3 --no-cpu-baseline --no-legs > $O/b12_bf16_c2_ch$ch.json 2> $O/b12_err.log
  python - $ch <<'PY'
import json, sys
try:
    a = json.load(open("gpurun_out/b12_bf16_c2_ch%s.json" % sys.argv[1]))
    print("bf16 c2 chunk", sys.argv[1], "value", round(a["value"]), "ms", round(a["ms_per_step"], 2), "launches", a["gpu_launches"], "graph_replays", a.get("graph_replays"), "host_ms", round(a.get("host_ms_per_step", 0), 2))
except Exception as ex:
    print("chunk", sys.argv[1], "failed", ex)
PY
done
for ch in 64 128 256 512; do
  timeout 300 python bench.py --precision tf32 --config c3 --chunk $ch --steps 4 --warmup 3 --no-cpu-baseline --no-legs > $O/b12_tf32_c3_ch$ch.json 2> $O/b12_err.log
  python - $ch <<'PY'
import json, sys
try:
    a = json.load(open("gpurun_out/b12_tf32_c3_ch%s.json" % sys.argv[1]))
    print("tf32 c3 chunk", sys.argv[1], "value", round(a["value"]), "ms", round(a["ms_per_step"], 2), "launches", a["gpu_launches"], "graph_replays", a.get("graph_replays"), "K1", round(a["roofline"]["frac"], 3), round(a["roofline"].get("frac_physical", 0), 3))
except Exception as ex:
    print("chunk", sys.argv[1], "failed", ex)
PY
done
DORPATCH_LANES=1 timeout 300 python bench.py --precision bf16 --config c2 --chunk 64 --steps 6 --warmup 3 --no-cpu-baseline --no-legs > $O/b12_bf16_c2_ch64_l1.json 2>> $O/b12_err.log
python -c "
import json; a = json.load(open('gpurun_out/b12_bf16_c2_ch64_l1.json')); print('bf16 c2 chunk 64 one lane', round(a['value']), round(a['ms_per_step'], 2))"
