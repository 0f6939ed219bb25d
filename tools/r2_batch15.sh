#!/bin/bash
# Round 2, multi-GPU batch (run under `gpurun --gpus N`): parity of the EOT-sharded step, then bench.py at the GPU counts given
#   gpurun --gpus 2 -- 'bash tools/r2_batch15.sh 2'        gpurun --gpus 8 -- 'bash tools/r2_batch15.sh 8 2 c5:4'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
NGPU=$(nvidia-smi -L | wc -l)
if [ "$1" = "2" ] && [ $# -eq 1 ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4
fi
port=29540
for spec in "$@"; do
  cfg=c3; n=$spec
  case $spec in c5:*) cfg=c5s0; n=${spec#c5:};; esac
  if [ $n -gt $NGPU ]; then echo "skip $spec: only $NGPU GPUs"; continue; fi
  port=$((port+1))
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --config $cfg --steps 6 --warmup 3 --no-legs --no-cpu-baseline > $O/b15_${cfg}_${n}gpu.json 2> $O/b15_${cfg}_${n}gpu.err
  tail -2 $O/b15_${cfg}_${n}gpu.err | cut -c1-300
  python - $O/b15_${cfg}_${n}gpu.json <<'PY'
import json, sys
try:
    a = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(a["config"]["workload"][:110])
    print("  n_gpus", a["n_gpus"], "value", round(a["value"]), a["dtype"], "ms/step", round(a["ms_per_step"], 2), "e2e", round(a["e2e"]["value"]), "host_ms", round(a.get("host_ms_per_step", 0), 2),
          "graph_replays", a.get("graph_replays"), "K1", round(a["roofline"]["frac"], 3), "launches", a["gpu_launches"])
except Exception as ex:
    print("parse failed", ex)
PY
done
