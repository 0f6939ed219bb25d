"""Multi-GPU parity check (run under torchrun, one process per GPU):
  1. one hot-loop step with the EOT axis sharded across ranks + exchange_shards  ==  the same step
     computed unsharded on every rank (losses / predictions / patch gradient);
  2. DorPatch.generate sharded  ~=  unsharded (same seeds), to the trajectory tolerance.
Prints DIST_CHECK OK on rank 0."""
import contextlib
import io
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dorpatch_b200 import masks as PM  # noqa: E402
from dorpatch_b200.attack import DorPatch, exchange_shards  # noqa: E402
from dorpatch_b200.resnetv2 import ResNetV2  # noqa: E402
from dorpatch_b200.utils import NormModel, get_normalize  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    os.environ["DORPATCH_PRECISION"] = "fp32"
    os.environ["DORPATCH_CHUNK"] = "16"
    H, B, S = 112, 2, 4 * world
    net = ResNetV2(seed=0)
    eng = net.engine(H, max_images=B)
    g = torch.Generator().manual_seed(5)
    x, m, p = (torch.rand(B, 3, H, H, generator=g).to(dev), torch.rand(B, 1, H, H, generator=g).to(dev),
               torch.rand(B, 3, H, H, generator=g).to(dev))
    table = PM.universe(H, 2)
    idx = np.random.RandomState(1).randint(0, 2520, (B, S))
    y = np.array([3, 700])
    s_loc = S // world
    G = torch.zeros_like(x)
    r = eng.attack_grad(x, m, p, PM.gather(table, idx[:, rank * s_loc:(rank + 1) * s_loc]), y, [True, False], 0.1, 4.0, 0, G,
                        S_total=S)
    loss_all, preds_all = exchange_shards(dist, G, r["loss_adv"], r["preds"])
    G2 = torch.zeros_like(x)
    r2 = eng.attack_grad(x, m, p, PM.gather(table, idx), y, [True, False], 0.1, 4.0, 0, G2, S_total=S)
    assert np.allclose(loss_all, r2["loss_adv"], atol=1e-5), np.abs(loss_all - r2["loss_adv"]).max()
    assert np.array_equal(preds_all, r2["preds"])
    a, b = G.double().flatten(), G2.double().flatten()
    cos = float(a @ b / (a.norm() * b.norm()))
    rel = float((a - b).norm() / b.norm())
    # different per-call batch sizes pick different conv/GEMM tilings -> last-bit differences -> a few
    # ReLU / max-pool gates flip (same effect as GPU-vs-CPU, measured rel 2.7e-3): bound, not bit-equality
    assert cos > 0.9999 and rel < 2e-2, (cos, rel)

    model = torch.nn.DataParallel(NormModel(net, get_normalize("imagenet", "resnetv2"))).cuda().eval()
    xr = torch.rand(1, 3, H, H, generator=torch.Generator().manual_seed(7)).to(dev)
    kw = dict(patch_budget=0.12, n_classes=1000, targeted=True, y=torch.tensor([17]).to(dev), max_iterations=4,
              sampling_size=4 * world, dropout=1)

    def run(shard):
        os.environ["DORPATCH_SHARD"] = "eot" if shard else "off"
        random.seed(1234); torch.manual_seed(1234); np.random.seed(1234)
        with contextlib.redirect_stdout(io.StringIO()):
            return DorPatch().generate(model, xr, save_dir=None, batch_id=0, **kw)
    m1, p1 = run(True)
    m0, p0 = run(False)
    iou = float(((m1 > .5) & (m0 > .5)).sum()) / max(float(((m1 > .5) | (m0 > .5)).sum()), 1)
    d = (p1 - p0).abs()
    frac = float((d > 0.0025).float().mean())
    # same statistical trajectory tolerance as tests/test_gpu_generate.py::_compare
    assert iou >= 0.8 and float(d.max()) <= 2 * 4 * 0.01 + 1e-6 and float(d.mean()) <= 0.0025 and frac <= 0.02 * 2 * 4, \
        (iou, float(d.max()), float(d.mean()), frac)
    dist.barrier()
    if rank == 0:
        print("DIST_CHECK OK world=%d: step loss max diff %.2e, grad cos %.9f rel %.2e; generate mask IoU %.3f pattern max diff %.4f"
              % (world, np.abs(loss_all - r2["loss_adv"]).max(), cos, rel, iou, float(d.max())))
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/dist_check_rank%s.err" % os.environ.get("RANK", "x"), "w") as f:
            traceback.print_exc(file=f)
        raise
