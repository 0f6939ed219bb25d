"""Oracle for the optional affine / colour EOT (dorpatch_b200/eot.py).  The reference has no such
transform, so this oracle is a plain PyTorch statement of the intended semantics
(F.affine_grid / F.grid_sample, align_corners=False, padding_mode='border', then the colour map) --
"parity unpinned" against the reference by construction; it pins the CUDA kernels
expand_affine_kernel / reduce_affine_kernel.  TEST INFRASTRUCTURE ONLY."""
import torch
import torch.nn.functional as F


def apply(adv_x, xf):
    """adv_x [B,3,H,W], xf [B,S,8] -> [B,S,3,H,W] (differentiable)."""
    B, _, H, W = adv_x.shape
    S = xf.shape[1]
    xf = torch.as_tensor(xf, dtype=torch.float32)
    theta = xf[..., :6].reshape(B * S, 2, 3)
    grid = F.affine_grid(theta, (B * S, 3, H, W), align_corners=False)
    src = adv_x[:, None].expand(B, S, 3, H, W).reshape(B * S, 3, H, W)
    warped = F.grid_sample(src, grid, mode="bilinear", padding_mode="border", align_corners=False)
    c = xf[..., 6].reshape(B * S, 1, 1, 1)
    b = xf[..., 7].reshape(B * S, 1, 1, 1)
    return torch.clamp(c * (warped - 0.5) + 0.5 + b, 0.0, 1.0).reshape(B, S, 3, H, W)
