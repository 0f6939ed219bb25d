"""K1 in-step variant (expand_kernel<FUSED=1>): achieved algorithmic GB/s over launch shapes and store paths, in one process.
    python tools/k1_step_sweep.py [bf16|tf32] [out.jsonl]
For every (samples per launch, EOT per image) shape of the bench configurations: the heuristic's choice, then the grid
tile rows x sample groups x store mode (dp_debug_k1_tuning).  One JSON line per point."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dorpatch_b200 import _lib, masks as PM
from dorpatch_b200.engine import Engine

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
out_path = sys.argv[2] if len(sys.argv) > 2 else None
IMG = 224
dev = torch.device("cuda", 0)
SHAPES = [(256, 16), (512, 16), (2048, 32), (128, 128)]          # (samples per launch, EOT per image): chunk, c2 whole, c3 whole, B=1
eng = Engine(img=IMG, precision=prec, chunk=8, max_images=128, device=0, autotune=False)
es = eng.elem_bytes
dt = torch.bfloat16 if es == 2 else torch.float32
table = PM.universe(IMG, 2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lines = []


def emit(d):
    print(json.dumps(d), flush=True)
    lines.append(d)


for n, S in SHAPES:
    B = max(1, 2 * n // S)
    g = torch.Generator().manual_seed(0)
    x, p = torch.rand(B, 3, IMG, IMG, generator=g).to(dev), torch.rand(B, 3, IMG, IMG, generator=g).to(dev)
    m = (torch.rand(B, 1, IMG, IMG, generator=g) > 0.9).float().to(dev)
    rects = PM.gather(table, np.stack([np.random.RandomState(b).choice(len(table), S, replace=False) for b in range(B)]))
    rd = torch.from_numpy(np.ascontiguousarray(rects.reshape(B * S, 4, 4), np.int16)).to(dev)
    eng.paste(x, m, p, 4.0)
    n_rot = max(2, int(400e6 // (n * IMG * IMG * eng.c_pad * es)) + 1)
    bufs = [torch.empty((n, IMG, IMG, eng.c_pad), dtype=dt, device=dev) for _ in range(n_rot)]
    starts = list(range(0, B * S - n + 1, n))

    def launch(i, rp):
        _lib.check(eng.lib.dp_expand_step_dev(eng.handle, C.c_void_p(x.data_ptr()), C.c_void_p(m.data_ptr()), C.c_void_p(p.data_ptr()), B, S, rp,
                                              starts[i % len(starts)], n, C.c_void_p(bufs[i % n_rot].data_ptr()), eng._stream()))

    def timeit(rp):
        for i in range(4):
            launch(i, rp)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0.record()
            for i in range(10):
                launch(i, rp)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 10)
        return float(np.median(ts))

    alg = n * IMG * IMG * 3 * es + (n // S if n >= S else 1) * 7 * IMG * IMG * 4
    grid = [(0, 0, 0)] + [(r, sg, mode) for mode in (0, 1) for r in (2, 4, 7, 8, 14, 16) for sg in (1, 2, 4) if sg * 8 <= S or sg == 1]
    for rows, sg, mode in grid:
        eng.lib.dp_debug_k1_tuning(rows, sg, mode)
        try:
            ms, ms_clean = timeit(C.c_void_p(rd.data_ptr())), timeit(None)
        except RuntimeError as ex:
            emit(dict(prec=prec, n=n, S=S, rows=rows, sg=sg, mode=mode, error=str(ex)[:80]))
            continue
        last = (C.c_int32 * 4)()
        eng.lib.dp_debug_k1_last(C.cast(last, C.c_void_p))
        emit(dict(prec=prec, n=n, S=S, rows=rows or "auto", sg=sg or "auto", mode=mode, shape=list(last), ms=round(ms, 4), gbs=round(alg / ms / 1e6),
                  frac=round(alg / ms / 1e6 / 6569.3, 3), clean_gbs=round(alg / ms_clean / 1e6)))
    del bufs
eng.lib.dp_debug_k1_tuning(0, 0, 0)
if out_path:
    with open(out_path, "w") as f:
        for d in lines:
            f.write(json.dumps(d) + "\n")
best = {}
for d in lines:
    if "frac" in d:
        k = (d["n"], d["S"])
        if k not in best or d["frac"] > best[k]["frac"]:
            best[k] = d
for k, d in sorted(best.items()):
    print("BEST", k, d)
