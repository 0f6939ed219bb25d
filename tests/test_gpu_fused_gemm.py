"""Whole-network check of the tcgen05 GroupNorm-in-the-prologue 1x1-convolution GEMM (kernels_gemm.cu,
DORPATCH_FUSED_GEMM=1) against the default engine path (cluster GroupNorm + cublasLt), bf16.

The operator itself is pinned against an fp32 restatement in tests/test_gpu_ops.py::test_tcgen05_gn_gemm_vs_fp32
(max |diff| <= one bf16 ulp of the output's range).  Here both paths round relu(gn(x)) to bf16 before the product and
accumulate in fp32; only the summation order differs, which this random-init network amplifies like any bf16
perturbation: moving a sample inside the batch on the DEFAULT path already changes the logits by 3 % relative L2
(precision; tools/additivity_diag.py).  Measured fused-vs-default on hardware (round 2): 2.8 % / 2.9 %.
The input gradient of this network is chaotic at bf16 (cosine 0.5 against fp32, tools/precision_diag.py), and the fused GEMM
re-rounds EVERY 1x1-convolution output (another summation order), so fused-vs-default is not a meaningful distance
(measured cosine 0.70).  The statement tested: the fused path is as close to the fp32 engine as the default bf16 path
is -- logits relative error <= 1.5 x the default path's, input-gradient cosine >= the default path's - 0.1."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(oracle_params, mode, H, N):
    """mode: 'fp32' (the reference arithmetic of this comparison), 'bf16' (default path), 'fused' (bf16 + tcgen05 GEMM)."""
    from dorpatch_b200.engine import Engine
    if mode == "fused":
        os.environ["DORPATCH_FUSED_GEMM"] = "1"
    else:
        os.environ.pop("DORPATCH_FUSED_GEMM", None)
    try:
        e = Engine(img=H, precision="fp32" if mode == "fp32" else "bf16", chunk=N, max_images=N, autotune=False)
    finally:
        os.environ.pop("DORPATCH_FUSED_GEMM", None)
    e.load_state_dict(oracle_params)
    g = torch.Generator().manual_seed(5)
    z = (torch.rand((N, 3, H, H), generator=g) * 2 - 1).to(DEV)
    dl = torch.randn((N, 1000), generator=g).to(DEV)
    logits, dz = e.net_forward_backward(z, dl)
    torch.cuda.synchronize()
    out = logits.float().cpu(), dz.float().cpu()
    e.close()
    return out


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float(a @ b / (a.norm() * b.norm()))


@pytest.mark.parametrize("H,N", [(224, 8), (224, 3)])      # M a multiple of 128 / ragged last tile
def test_fused_gn_gemm_is_as_accurate_as_the_default_bf16_path(oracle_params, H, N):
    lt, gt = _run(oracle_params, "fp32", H, N)              # fp32 engine (== oracle to 4e-6, tests/test_gpu_fullsize.py)
    l0, g0 = _run(oracle_params, "bf16", H, N)
    l1, g1 = _run(oracle_params, "fused", H, N)
    rel0, rel1 = float((l0 - lt).norm() / lt.norm()), float((l1 - lt).norm() / lt.norm())
    cos0, cos1 = _cos(g0, gt), _cos(g1, gt)
    print("\nlogits rel L2 vs fp32: default bf16 %.3e, fused %.3e; fused vs default %.3e" % (rel0, rel1, float((l1 - l0).norm() / l0.norm())))
    print("input-gradient cosine vs fp32: default bf16 %.4f, fused %.4f; fused vs default %.4f" % (cos0, cos1, _cos(g1, g0)))
    assert np.isfinite(l1.numpy()).all() and np.isfinite(g1.numpy()).all()
    assert rel1 <= 1.5 * rel0 + 1e-3, (rel1, rel0)
    assert cos1 >= cos0 - 0.1, (cos1, cos0)
