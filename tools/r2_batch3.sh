#!/bin/bash
# Round 2, GPU batch 3: phase trace of the GN v2 forward kernel, round-1 membench step variants, new op-level / 224-px tests, new bench.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o $O/gnbench tools/gnbench.cu -Ldorpatch_b200/lib -ldorpatch -Xlinker -rpath,$PWD/dorpatch_b200/lib 2> $O/gnbench_build.log || { cat $O/gnbench_build.log; exit 1; }
for c in 64 256 512 1024; do
  GNBENCH_TRACE=1 DORPATCH_GN=v2 timeout 300 $O/gnbench 256 $c 2>&1 | grep -v NEG | tee -a $O/r2_gnbench_trace.log
done
rm -f $O/gnbench
timeout 60 nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o $O/membench tools/membench.cu > $O/r2_membench.log 2>&1 && timeout 200 $O/membench >> $O/r2_membench.log 2>&1
cat $O/r2_membench.log | tail -60; rm -f $O/membench
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -70 > $O/r2_b3_pytest.log
cat $O/r2_b3_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_b3_bench.json 2> $O/r2_b3_bench.err
tail -5 $O/r2_b3_bench.err
python - <<'PY'
import json
try:
    a = json.load(open("gpurun_out/r2_b3_bench.json"))
    print("value", round(a["value"]), a["dtype"], "ms", round(a["ms_per_step"], 1), "e2e", round(a["e2e"]["value"]), "launches", a["gpu_launches"])
    print("roofline", {k: a["roofline"][k] for k in ("frac", "ms", "ms_single_launch_event_pair", "samples_per_launch", "unoccluded_gbs")})
    print({n: (v["ms"], v.get("frac_of_hbm_peak", v.get("frac_of_bf16_peak"))) for n, v in list(a["kernels"].items())[:12]})
    print("scan", a.get("scan"), "pc", a.get("patchcleanser_eval"))
    for k, v in a.get("legs", {}).items():
        print(k, {q: (round(w, 1) if isinstance(w, float) else w) for q, w in v.items() if q not in ("kernels", "config")})
    k = a.get("legs", {}).get("bf16_c2", {}).get("kernels")
    if k: print("bf16_c2 kernels", {n: (v["ms"], v.get("frac_of_hbm_peak", v.get("frac_of_bf16_peak"))) for n, v in list(k.items())[:8]})
except Exception as ex:
    print("bench parse failed", ex)
PY
