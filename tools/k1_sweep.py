"""Time the standalone K1 launch (dp_expand_dev) on the bench workload under the current environment
(DORPATCH_K1_ROWS / DORPATCH_K1_SG ...), one JSON line.  The launch tunables are read once per
process, so a sweep runs this script once per setting:

    for sg in 1 2; do DORPATCH_K1_SG=$sg python tools/k1_sweep.py; done
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dorpatch_b200 import _lib, masks as PM          # noqa: E402
from dorpatch_b200.engine import Engine               # noqa: E402


def main():
    B, S, IMG = int(os.environ.get("K1_B", 32)), int(os.environ.get("K1_S", 16)), 224
    prec = os.environ.get("K1_PRECISION", "bf16")
    dev = torch.device("cuda:0")
    eng = Engine(img=IMG, n_classes=1000, precision=prec, chunk=64, max_images=B, autotune=False)
    table = PM.universe(IMG, 2)
    n_mask = table.shape[0]
    g = torch.Generator().manual_seed(0)
    x = torch.rand((B, 3, IMG, IMG), generator=g).to(dev)
    idx = np.stack([np.random.RandomState(b).choice(n_mask, S, replace=False) for b in range(B)])
    rects = np.ascontiguousarray(PM.gather(table, idx).reshape(B * S, 4, 4), np.int16)
    rd = torch.from_numpy(rects).to(dev)
    es = eng.elem_bytes
    buf = torch.empty((B * S, IMG, IMG, eng.c_pad), dtype=torch.bfloat16 if es == 2 else torch.float32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run(rp, reps):
        ts = []
        for i in range(reps + 3):
            e0.record()
            _lib.check(eng.lib.dp_expand_dev(eng.handle, C.c_void_p(x.data_ptr()), B, S, rp, C.c_void_p(buf.data_ptr()), eng._stream()))
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    ms = run(C.c_void_p(rd.data_ptr()), 30)
    ms0 = run(None, 10)
    alg = B * S * IMG * IMG * 3 * es + B * 3 * IMG * IMG * 4
    env = {k: v for k, v in os.environ.items() if k.startswith("DORPATCH_K1")}
    print(json.dumps({"env": env, "ms": ms, "gbs": alg / ms / 1e6, "ms_unoccluded": ms0, "gbs_unoccluded": alg / ms0 / 1e6}))


if __name__ == "__main__":
    main()
