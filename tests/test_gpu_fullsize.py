"""GPU-vs-oracle parity AT 224 px (BASELINE.json's image size): the classifier forward / backward-to-input and one whole
hot-loop step, fp32 engine against the torch-CPU oracle.  The 112-px tests in test_gpu_kernels.py never reach the
224-only kernel variants (GroupNorm clusters of 8 / 16 CTAs with 512-thread CTAs for the 56x56 layers, the dy-streamed
backward, K1 tiles of a full-width row); these do.  Costs a few seconds of CPU for the oracle side."""
import numpy as np
import pytest
import torch

from oracle import attack as OA
from oracle import masks as OM
from oracle import resnetv2 as OR

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
H = 224


def _rand(shape, seed):
    return torch.rand(shape, generator=torch.Generator().manual_seed(seed))


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def _rects_for(idx):
    table = OM.rects_to_array(OM.universe_rects(H, 2))
    out = np.zeros(np.asarray(idx).shape + (4, 4), np.int16)
    out[..., 0:2, :] = table[np.asarray(idx)]
    return out


def test_classifier_forward_backward_224_fp32(engine_factory, oracle_params):
    """ResNetV2-50 logits and d/d(input) at 224 px, N = 2, fp32 engine vs torch-CPU autograd on the oracle restatement
    (utils.py:77-78 -> timm resnetv2_50x1_bit): |dlogit| <= 2e-3, gradient cosine >= 0.999."""
    N = 2
    e = engine_factory(img=H, precision="fp32", chunk=6, max_images=2)
    z = (_rand((N, 3, H, H), 11) - 0.5) * 2
    dl = torch.zeros(N, 1000)
    dl[torch.arange(N), torch.tensor([3, 999])] = 1.0
    dl[torch.arange(N), torch.tensor([7, 0])] = -1.0
    logits, dz = e.net_forward_backward(z.to(DEV), dl.to(DEV))
    torch.cuda.synchronize()
    zr = z.clone().requires_grad_(True)
    ref = OR.forward_normalized(oracle_params, zr)
    (ref * dl).sum().backward()
    err = (logits.cpu() - ref.detach()).abs().max().item()
    cos = _cos(dz.cpu(), zr.grad)
    print("224 px fp32: logit err", err, "grad cos", cos)
    assert err <= 2e-3, err
    assert cos >= 0.999, cos


def test_attack_step_224_matches_oracle(engine_factory, oracle_net):
    """One hot-loop iteration (attack.py:184-247 + :332-342) at 224 px, fp32 engine, B = 2 x S = 3 in one chunk of 6:
    loss_adv 2e-3 abs, structural loss rtol 1e-4, pattern-gradient cosine >= 0.999 and sign agreement >= 99 % where
    |g| > 1e-3 of its max, exact sign step."""
    B, S = 2, 3
    e = engine_factory(img=H, precision="fp32", chunk=6, max_images=2)
    x, m, p = _rand((B, 3, H, H), 31), (_rand((B, 1, H, H), 32) > 0.8).float(), _rand((B, 3, H, H), 33)
    y = torch.tensor([17, 400])
    targeted = [True, False]
    idx = np.random.RandomState(3).randint(0, 2520, (B, S))
    structured = [1e-3, 5e-4]
    xd, md, pd = x.to(DEV), m.to(DEV).clone(), p.to(DEV).clone()
    G = torch.zeros(B, 3, H, H, device=DEV)
    r = e.attack_grad(xd, md, pd, _rects_for(idx), y.numpy(), targeted, 0.1, 4.0, 1, G)
    gp = torch.zeros_like(pd)
    lr = np.array([0.01, 0.02], np.float32)
    e.attack_update(xd, md, pd, G, lr, structured, None, 1e-3, 1, grad_pattern_out=gp)
    torch.cuda.synchronize()
    uni = torch.from_numpy(OM.rects_to_bool(OM.universe_rects(H, 2), H))
    lvx = OA.local_variance(x)[0].mean(1)
    o = OA.step_losses_and_grads(oracle_net, x, m, p, y, idx, uni, targeted, 1000, 0.1, structured, 1e-3, [1e-5, 1e-5], 1, 4.0, lvx)
    assert np.abs(r["loss_adv"] - o["loss_adv"].numpy()).max() <= 2e-3
    assert np.allclose(r["loss_struc"], o["loss_struc"].numpy(), rtol=1e-4, atol=1e-6)
    got, ref = gp.cpu(), o["grad_pattern"]
    cos = _cos(got, ref)
    big = ref.abs() > 1e-3 * ref.abs().max()
    agree = (got.sign()[big] == ref.sign()[big]).float().mean().item()
    print("224 px step: grad cos", cos, "sign agreement", agree)
    assert cos >= 0.999 and agree >= 0.99, (cos, agree)
    pref = (p - torch.from_numpy(lr)[:, None, None, None] * got.sign()).clamp(0, 1)
    assert torch.equal(pd.cpu(), pref)
