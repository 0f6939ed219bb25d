"""Occlusion-mask geometry of the PatchCleanser mask window, as rectangle tables.

Every mask the reference materialises as a [1,H,W] bool tensor
(/root/reference/defenses/PatchCleanser.py:8-59, /root/reference/attack.py:25-31,83-85) is
"keep everything except <= 2 axis-aligned rectangles"; the native K1 kernel synthesises the
occlusion from the rectangles (int16 r0,r1,c0,c1 -- rows [r0,r1) x cols [c0,c1) become 0.5),
so the 126 MB bool universe is never built.
"""
import math

import numpy as np

DROPOUT_SIZES = (0.015, 0.03, 0.06, 0.12)
MASKS_PER_AXIS = 6


class WindowGeometry:
    """mask_size / stride / window_size of one patch ratio (PatchCleanser.py:11-16)."""

    def __init__(self, img_size, patch_ratio, n_patch=1):
        self.img_size = int(img_size)
        self.mask_size = math.floor(math.sqrt(img_size ** 2 * patch_ratio / n_patch))
        self.stride = int(np.ceil((img_size - self.mask_size + 1) / MASKS_PER_AXIS))
        self.window_size = self.mask_size + self.stride - 1

    def single_rects(self):
        """int16 [36,4]: mask k = i*6+j starts at row stride*i, col stride*j."""
        k = np.arange(MASKS_PER_AXIS)
        lo = self.stride * k
        hi = np.minimum(self.img_size, lo + self.window_size)
        r = np.zeros((MASKS_PER_AXIS, MASKS_PER_AXIS, 4), np.int16)
        r[:, :, 0], r[:, :, 1] = lo[:, None], hi[:, None]
        r[:, :, 2], r[:, :, 3] = lo[None, :], hi[None, :]
        return r.reshape(-1, 4)


def pair_index():
    """(a,b), a<b, row-major over the 36x36 upper triangle (PatchCleanser.py:21-29)."""
    a, b = np.triu_indices(MASKS_PER_AXIS ** 2, k=1)
    return a, b


def mask_set(img_size, patch_ratio, dropout):
    """int16 [n,2,4] rectangle pairs of `mask_set` (dropout 1) / `double_mask_set` (dropout 2)."""
    s = WindowGeometry(img_size, patch_ratio).single_rects()
    if dropout == 1:
        out = np.zeros((len(s), 2, 4), np.int16)
        out[:, 0] = s
        return out
    if dropout == 2:
        a, b = pair_index()
        return np.stack([s[a], s[b]], axis=1)
    raise ValueError("dropout must be 1 or 2, got %r" % (dropout,))


def universe(img_size, dropout, sizes=DROPOUT_SIZES):
    """The attack's mask universe (attack.py:83-85): [n_mask,2,4] int16."""
    return np.concatenate([mask_set(img_size, r, dropout) for r in sizes], axis=0)


def gather(table, idx, idx_dual=None):
    """Rectangles of the sampled masks: table [n,2,4], idx [...] -> [...,4,4] int16
    (second pair from idx_dual, or empty)."""
    idx = np.asarray(idx)
    out = np.zeros(idx.shape + (4, 4), np.int16)
    out[..., 0:2, :] = table[idx]
    if idx_dual is not None:
        out[..., 2:4, :] = table[np.asarray(idx_dual)]
    return out


def to_bool(table, img_size):
    """Materialise [n,1,H,W] bool masks (True = keep) -- only for API compatibility
    (MaskWindow.mask_set) and tests; the engine never uses it."""
    n = table.shape[0]
    m = np.ones((n, 1, img_size, img_size), dtype=bool)
    for k in range(n):
        for r0, r1, c0, c1 in table[k].reshape(-1, 4):
            m[k, 0, r0:r1, c0:c1] = False
    return m
