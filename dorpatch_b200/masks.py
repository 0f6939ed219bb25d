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


def from_bool(masks):
    """Inverse of to_bool for API compatibility: bool keep-masks [n,1,H,W] (torch or numpy) -> rectangle table
    [n,2,4] int16.  Each mask's occluded set must be a union of <= 2 axis-aligned rectangles, as the reference
    builds them (PatchCleanser.py:51-58: one rectangle per single mask, the union of two per double mask); anything
    else raises -- the engine has no dense-mask path."""
    m = masks.detach().cpu().numpy() if hasattr(masks, "detach") else np.asarray(masks)
    m = m.reshape(m.shape[0], m.shape[-2], m.shape[-1]).astype(bool)
    out = np.zeros((m.shape[0], 2, 4), np.int16)
    for k in range(m.shape[0]):
        out[k] = _decompose(~m[k], k)
    return out


def _decompose(occ, k=0):
    """occ (bool [H,W]) as a union of <= 2 rectangles.  Work on the coarse grid of rows / columns where the pattern
    changes (<= 5 x 5 cells): the first rectangle starts at the first occluded cell; try each of its extents, the
    second rectangle is then the bounding box of what is left, valid if it lies inside the occluded set."""
    H, W = occ.shape
    out = np.zeros((2, 4), np.int16)
    if not occ.any():
        return out
    rb = [0] + [i for i in range(1, H) if not np.array_equal(occ[i], occ[i - 1])] + [H]
    cb = [0] + [j for j in range(1, W) if not np.array_equal(occ[:, j], occ[:, j - 1])] + [W]
    g = occ[np.ix_(rb[:-1], cb[:-1])]
    a0 = int(np.argmax(g.any(1)))
    b0 = int(np.argmax(g[a0]))
    for a1 in range(a0 + 1, g.shape[0] + 1):
        for b1 in range(b0 + 1, g.shape[1] + 1):
            if not g[a0:a1, b0:b1].all():
                continue
            rest = g.copy()
            rest[a0:a1, b0:b1] = False
            if not rest.any():
                out[0] = (rb[a0], rb[a1], cb[b0], cb[b1])
                return out
            rr, cc = np.nonzero(rest)
            # the second rectangle may overlap the first: grow the remainder's bounding box over occluded cells
            for c0 in range(int(rr.min()), -1, -1):
                for d0 in range(int(cc.min()), -1, -1):
                    c1, d1 = int(rr.max()) + 1, int(cc.max()) + 1
                    if not g[c0:c1, d0:d1].all():
                        continue
                    u = np.zeros_like(g)
                    u[a0:a1, b0:b1] = True
                    u[c0:c1, d0:d1] = True
                    if np.array_equal(u, g):
                        out[0] = (rb[a0], rb[a1], cb[b0], cb[b1])
                        out[1] = (rb[c0], rb[c1], cb[d0], cb[d1])
                        return out
    raise NotImplementedError("mask %d is not a union of <= 2 axis-aligned rectangles" % k)
