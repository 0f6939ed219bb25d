"""Diagnostic: how additive over EOT shards / how reproducible is dp_attack_grad at the full bench size?
Prints one JSON line per engine configuration (no assertions)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dorpatch_b200.engine import Engine          # noqa: E402
from oracle import masks as OM                   # noqa: E402  (diagnostic tool, not the product path)
from oracle import resnetv2 as R                 # noqa: E402

DEV = "cuda:0"
H, B, S = 224, 32, 16


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def main():
    params = R.random_init(seed=0, affine_jitter=0.1)
    g = lambda shape, seed: torch.rand(shape, generator=torch.Generator().manual_seed(seed))
    x, m, p = g((B, 3, H, H), 1).to(DEV), (g((B, 1, H, H), 2) * 0.2).to(DEV), g((B, 3, H, H), 3).to(DEV)
    table = OM.rects_to_array(OM.universe_rects(H, 2))
    idx = np.stack([np.random.RandomState(100 + b).choice(table.shape[0], S, replace=False) for b in range(B)])
    rects = np.zeros((B, S, 4, 4), np.int16)
    rects[:, :, 0:2, :] = table[idx]
    for prec, autotune, chunk in (("bf16", False, 256), ("bf16", True, 256), ("bf16", True, 512), ("fp32", False, 128)):
        e = Engine(img=H, precision=prec, chunk=chunk, max_images=B, autotune=autotune)
        e.load_state_dict(params)
        y = e.predict(x).astype(np.int64)

        def grad(r):
            G = torch.zeros(B, 3, H, H, device=DEV)
            out = e.attack_grad(x, m, p, r, y, [False] * B, 0.1, 4.0, 1, G, S_total=S)
            torch.cuda.synchronize()
            return G.cpu(), out["loss_adv"].copy(), out["preds"].copy()

        G, l, pr = grad(rects)
        G2, l2, _ = grad(rects)
        Ga, la, pa = grad(rects[:, :S // 2])
        Gb, lb, pb = grad(rects[:, S // 2:])
        lcat = np.concatenate([la, lb], 1)
        per_img = [cos((Ga + Gb)[b], G[b]) for b in range(B)]
        print(json.dumps({"precision": prec, "autotune": autotune, "chunk": chunk,
                          "rerun_cos": cos(G2, G), "rerun_rel": rel(G2, G), "rerun_loss_maxdiff": float(np.abs(l2 - l).max()),
                          "add_cos": cos(Ga + Gb, G), "add_rel": rel(Ga + Gb, G), "add_loss_maxdiff": float(np.abs(lcat - l).max()),
                          "add_cos_per_image_min": min(per_img), "preds_equal": float((np.concatenate([pa, pb], 1) == pr).mean()),
                          "loss_range": [float(l.min()), float(l.max())]}), flush=True)
        e.close()


if __name__ == "__main__":
    main()
