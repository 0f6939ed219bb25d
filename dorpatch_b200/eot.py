"""Optional affine / colour EOT transforms.

The north-star names "EOT affine/colour transforms"; the reference contains none (only an unused
`transforms=None` hook on collect_failure, attack.py:384,395-396).  They are therefore an opt-in
extension, default OFF (strength 0 => the engine takes the reference's exact path).  A transform is
8 floats per EOT sample: a 2x3 affine theta in torch `affine_grid` convention (normalised
coordinates, align_corners=False, border padding) followed by v' = clamp(contrast*(v-0.5)+0.5+
brightness, 0, 1), applied to the pasted image before the occlusion mask.
"""
import numpy as np

MAX_ROT = np.pi / 6      # at strength 1: +-30 degrees
MAX_SCALE = 0.10         # +-10 %
MAX_SHIFT = 0.10         # +-5 % of the width (normalised coordinates span 2)
MAX_CONTRAST = 0.20
MAX_BRIGHT = 0.10


def sample(rng, B, S, affine=0.0, colour=0.0):
    """[B,S,8] float32 transforms drawn from `rng` (a numpy RandomState separate from the global
    stream, so the occlusion sampling of attack.py:193-204 is not disturbed)."""
    u = rng.uniform(-1.0, 1.0, size=(6, B, S))
    ang = u[0] * affine * MAX_ROT
    sc = 1.0 + u[1] * affine * MAX_SCALE
    out = np.zeros((B, S, 8), np.float32)
    out[..., 0], out[..., 1], out[..., 2] = sc * np.cos(ang), -sc * np.sin(ang), u[2] * affine * MAX_SHIFT
    out[..., 3], out[..., 4], out[..., 5] = sc * np.sin(ang), sc * np.cos(ang), u[3] * affine * MAX_SHIFT
    out[..., 6] = 1.0 + u[4] * colour * MAX_CONTRAST
    out[..., 7] = u[5] * colour * MAX_BRIGHT
    return out
