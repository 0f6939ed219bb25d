"""Size-independent properties of the hot path at BASELINE.json's full configuration
(configs[1]: 32 images x 16 EOT occlusion samples, 224x224; fp32 engine with chunk 128 and the
bench's bf16 engine with chunk 256 x 2 lanes).

The CPU oracle needs minutes per step at this size, so parity is checked through identities the
algorithm guarantees (attack.py:184-247: the step's gradient is a plain sum over EOT samples of
per-sample gradients, each a function of that sample's image and rectangles only):

  * additivity over EOT shards -- the identity the multi-GPU path relies on;
  * invariance under a permutation of the samples of each image;
  * a fully occluded sample sees the constant 0.5 image: image-independent logits, exactly zero gradient;
  * K1 at full size against a direct numpy construction, bit for bit.

Tolerances.  fp32 engine: a sample's result does not depend on its position in the batch, only
the fp32 accumulation order of G changes -> cosine > 1 - 1e-6, relative L2 < 1e-5, losses to 1e-5
(measured: 7e-8 relative).  bf16 engine: the library GEMMs' tile / split order depends on the row
position, so moving a sample changes single bf16 roundings, which this random-init network
amplifies (precision) -- measured with tools/additivity_diag.py at this size: cosine
0.99955, relative L2 0.030, losses to 5.5e-4, bit-identical run to run -> bars cosine > 0.995,
relative L2 < 0.1, losses to 5e-3.
"""
import numpy as np
import pytest
import torch

from oracle import masks as OM

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
H, B, S = 224, 32, 16


def _rand(shape, seed):
    return torch.rand(shape, generator=torch.Generator().manual_seed(seed))


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


BARS = {"fp32": dict(cos=1 - 1e-6, rel=1e-5, loss=1e-5), "bf16": dict(cos=0.995, rel=0.1, loss=5e-3)}


def _close(a, b, bars):
    return _cos(a, b) > bars["cos"] and _rel(a, b) < bars["rel"]


@pytest.fixture(scope="module", params=[("fp32", 128), ("bf16", 256)], ids=["fp32", "bf16"])
def full(request, oracle_params):
    from dorpatch_b200.engine import Engine
    precision, chunk = request.param
    e = Engine(img=H, precision=precision, chunk=chunk, max_images=B, autotune=False)
    e.load_state_dict(oracle_params)
    x, m, p = _rand((B, 3, H, H), 1).to(DEV), (_rand((B, 1, H, H), 2) * 0.2).to(DEV), _rand((B, 3, H, H), 3).to(DEV)
    table = OM.rects_to_array(OM.universe_rects(H, 2))
    idx = np.stack([np.random.RandomState(100 + b).choice(table.shape[0], S, replace=False) for b in range(B)])
    rects = np.zeros((B, S, 4, 4), np.int16)
    rects[:, :, 0:2, :] = table[idx]
    y = e.predict(x).astype(np.int64)
    yield e, x, m, p, rects, y, BARS[precision]
    e.close()


def _grad(e, x, m, p, rects, y, S_total, stage=1):
    G = torch.zeros(x.shape[0], 3, H, H, device=DEV)
    r = e.attack_grad(x, m, p, rects, y, [False] * x.shape[0], 0.1, 4.0, stage, G, S_total=S_total)
    torch.cuda.synchronize()
    return G.cpu(), r["loss_adv"].copy(), r["preds"].copy()


def test_gradient_is_additive_over_eot_shards(full):
    e, x, m, p, rects, y, bars = full
    G, loss, preds = _grad(e, x, m, p, rects, y, S)
    Ga, la, pa = _grad(e, x, m, p, rects[:, :S // 2], y, S)
    Gb, lb, pb = _grad(e, x, m, p, rects[:, S // 2:], y, S)
    assert G.abs().max() > 0
    assert _close(Ga + Gb, G, bars)
    assert np.allclose(np.concatenate([la, lb], 1), loss, atol=bars["loss"])
    assert (np.concatenate([pa, pb], 1) == preds).mean() > 0.99


def test_gradient_is_invariant_under_sample_permutation(full):
    e, x, m, p, rects, y, bars = full
    G, loss, _ = _grad(e, x, m, p, rects, y, S)
    perm = np.random.RandomState(7).permutation(S)
    Gp, lp, _ = _grad(e, x, m, p, rects[:, perm], y, S)
    assert _close(Gp, G, bars)
    assert np.allclose(lp, loss[:, perm], atol=bars["loss"])


def test_fully_occluded_samples_are_constant_and_gradient_free(full):
    e, x, m, p, rects, y, bars = full
    r2 = rects.copy()
    r2[:, 0] = 0
    r2[:, 0, 0] = (0, H, 0, H)                 # sample 0 of every image: one rectangle over the whole image
    r2[0] = r2[0, 0]                           # ... and every sample of image 0
    G, loss, preds = _grad(e, x, m, p, r2, y, S)
    # a fully occluded sample is the same constant image whatever b: one prediction, and with equal labels one loss
    occluded = np.concatenate([preds[0], preds[1:, 0]])
    assert len(set(occluded.tolist())) == 1
    assert np.ptp(loss[0]) <= bars["loss"]
    same = y == y[0]
    assert np.ptp(loss[same, 0]) <= bars["loss"]
    # ... and it passes no gradient to the patch: image 0 (all samples occluded) gets exactly zero
    assert float(G[0].abs().max()) == 0.0 and float(G[1:].abs().max()) > 0.0


def test_k1_full_size_bit_exact(full):
    e, x, m, p, rects, y, bars = full
    out = e.expand(x, S, rects)                # [B*S, H, H, c_pad] engine dtype
    torch.cuda.synchronize()
    got = out[..., :3].float().cpu().reshape(B, S, H, H, 3)
    assert float(out[..., 3:].abs().sum()) == 0.0      # channel padding (if any) is zero
    want = ((x.cpu() - 0.5) * 2.0).to(out.dtype).float().permute(0, 2, 3, 1)[:, None].repeat(1, S, 1, 1, 1)
    for b in range(B):
        for s in range(S):
            for k in range(4):
                r0, r1, c0, c1 = (int(v) for v in rects[b, s, k])
                if c1 > c0:
                    want[b, s, r0:r1, c0:c1] = 0.0
    assert torch.equal(got, want)
