"""K1 in-step variant (expand_kernel<FUSED=1>, one classifier chunk per launch): achieved algorithmic GB/s for the launch
shape selected by DORPATCH_K1_ROWS / DORPATCH_K1_SG (unset = the wave-efficiency heuristic).  One JSON line per call.
    DORPATCH_K1_ROWS=7 DORPATCH_K1_SG=1 python tools/k1_step_sweep.py bf16 256 16
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dorpatch_b200 import _lib, masks as PM
from dorpatch_b200.engine import Engine

prec, n_chunk, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
B = max(1, 2 * n_chunk // S)
IMG = 224
dev = torch.device("cuda", 0)
eng = Engine(img=IMG, precision=prec, chunk=8, max_images=B, device=0, autotune=False)
g = torch.Generator().manual_seed(0)
x, p = torch.rand(B, 3, IMG, IMG, generator=g).to(dev), torch.rand(B, 3, IMG, IMG, generator=g).to(dev)
m = (torch.rand(B, 1, IMG, IMG, generator=g) > 0.9).float().to(dev)
table = PM.universe(IMG, 2)
rects = PM.gather(table, np.stack([np.random.RandomState(b).choice(len(table), S, replace=False) for b in range(B)]))
rd = torch.from_numpy(np.ascontiguousarray(rects.reshape(B * S, 4, 4), np.int16)).to(dev)
es = eng.elem_bytes
dt = torch.bfloat16 if es == 2 else torch.float32
n_rot = max(2, int(400e6 // (n_chunk * IMG * IMG * eng.c_pad * es)) + 1)
bufs = [torch.empty((n_chunk, IMG, IMG, eng.c_pad), dtype=dt, device=dev) for _ in range(n_rot)]
starts = list(range(0, B * S - n_chunk + 1, n_chunk))


def launch(i, rp):
    _lib.check(eng.lib.dp_expand_step_dev(eng.handle, C.c_void_p(x.data_ptr()), C.c_void_p(m.data_ptr()), C.c_void_p(p.data_ptr()), B, S, rp,
                                          starts[i % len(starts)], n_chunk, C.c_void_p(bufs[i % n_rot].data_ptr()), eng._stream()))


def timeit(rp):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(6):
        launch(i, rp)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0.record()
        for i in range(10):
            launch(i, rp)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 10)
    return float(np.median(ts))


alg = n_chunk * IMG * IMG * 3 * es + (n_chunk // S) * 7 * IMG * IMG * 4
ms, ms_clean = timeit(C.c_void_p(rd.data_ptr())), timeit(None)
print(json.dumps(dict(prec=prec, n=n_chunk, S=S, rows=os.environ.get("DORPATCH_K1_ROWS", "auto"), sg=os.environ.get("DORPATCH_K1_SG", "auto"),
                      ms=round(ms, 4), gbs=round(alg / ms / 1e6), frac=round(alg / ms / 1e6 / 6569.3, 3), clean_gbs=round(alg / ms_clean / 1e6))))
