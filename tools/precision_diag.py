"""How far do reduced-precision GPU runs of the SAME network drift from the fp32 CPU oracle?
Compares (a) PyTorch-GPU fp32 / TF32 / bf16-autocast autograd and (b) the native engine in
fp32 / tf32 / bf16 against torch-CPU fp32: logit error and input-gradient cosine.  Answers
whether a low gradient cosine at tf32/bf16 is inherent to the (random-init, chaotic) network
or an engine defect."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import resnetv2 as OR
from dorpatch_b200.engine import Engine


def cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(a @ b / (a.norm() * b.norm()))


def main():
    H, N = int(os.environ.get("DIAG_IMG", "224")), 4
    res = {}
    for jitter in (0.1,):
        params = OR.random_init(seed=0, affine_jitter=jitter)
        g = torch.Generator().manual_seed(11)
        z = (torch.rand(N, 3, H, H, generator=g) - 0.5) * 2
        dl = torch.zeros(N, 1000)
        dl[torch.arange(N), torch.tensor([3, 500, 999, 17])] = 1.0
        dl[torch.arange(N), torch.tensor([7, 1, 0, 400])] = -1.0
        zr = z.clone().requires_grad_(True)
        ref = OR.forward_normalized(params, zr)
        (ref * dl).sum().backward()
        gref = zr.grad
        pg = {k: v.cuda() for k, v in params.items()}
        for name in ("fp32", "tf32", "bf16"):
            torch.backends.cudnn.allow_tf32 = name != "fp32"
            torch.backends.cuda.matmul.allow_tf32 = name != "fp32"
            zc = z.cuda().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(name == "bf16")):
                out = OR.forward_normalized(pg, zc)
            (out.float() * dl.cuda()).sum().backward()
            res["torch_gpu_" + name] = dict(logit_err=float((out.float().cpu() - ref.detach()).abs().max()),
                                            grad_cos=cos(zc.grad, gref))
        for name in ("fp32", "tf32", "bf16"):
            e = Engine(img=H, precision=name, chunk=4, max_images=1, autotune=False)
            e.load_state_dict(params)
            lg, dz = e.net_forward_backward(z.cuda(), dl.cuda())
            torch.cuda.synchronize()
            res["engine_" + name] = dict(logit_err=float((lg.cpu() - ref.detach()).abs().max()), grad_cos=cos(dz, gref))
            e.close()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
