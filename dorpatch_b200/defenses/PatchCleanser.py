"""PatchCleanser defense (evaluation of the generated patch) on the native forward engine.

Mirrors /root/reference/defenses/PatchCleanser.py: MaskWindow :6-59, PatchCleanser :62-118
(robust_predict :68-97, mask :99-100, robustness_certificate :102-112), PatchCleanserRecord
:121-126, PatchCleanserResult :129-134.  Masks are rectangle tables (dorpatch_b200.masks);
the bool mask tensors of the reference are materialised lazily, only if someone reads
``mask_set`` / ``double_mask_set``.
"""
import math

import numpy as np
import torch

from .. import masks as _masks
from ..utils import unwrap_native


class MaskWindow(object):
    def __init__(self, img_size, patch_ratio=0.03, n_patch=1):
        if n_patch != 1:
            raise NotImplementedError("n_patch=%r (the reference's n_patch=2 path is unused by DorPatch)" % n_patch)
        g = _masks.WindowGeometry(img_size, patch_ratio, n_patch)
        self.img_size, self.n_patch = img_size, n_patch
        self.mask_size, self.stride, self.window_size = g.mask_size, g.stride, g.window_size
        self.num_mask_per_axis = _masks.MASKS_PER_AXIS
        self.mask_rects = _masks.mask_set(img_size, patch_ratio, 1)           # [36,2,4]
        self.double_mask_rects = _masks.mask_set(img_size, patch_ratio, 2)    # [630,2,4]
        self._bool = {}
        print("mask size: %d, window size: %d, stride: %d" % (self.mask_size, self.window_size, self.stride))

    def _materialise(self, name, table):
        if name not in self._bool:
            t = torch.from_numpy(_masks.to_bool(table, self.img_size))
            self._bool[name] = t.cuda() if torch.cuda.is_available() else t
        return self._bool[name]

    @property
    def mask_set(self):
        return self._materialise("single", self.mask_rects)

    @property
    def double_mask_set(self):
        return self._materialise("double", self.double_mask_rects)

    @property
    def reverse_mask_set(self):
        return ~self.mask_set


class PatchCleanser(object):
    def __init__(self, mask_window, model, result=None):
        self.mask_window = mask_window
        self.model = model
        self.result = result
        self._net = unwrap_native(model)

    def _engine(self, img):
        return self._net.engine(img.shape[-1], max_images=1)

    def _predict(self, img, rects):
        """argmax predictions of the classifier on occlude(img, rects[k]); rects [n,4,4]."""
        eng = self._engine(img)
        x = img.reshape(1, 3, img.shape[-2], img.shape[-1]).contiguous().float()
        return eng.predict(x, rects.shape[0], rects)

    def robust_predict(self, img, certify=False):
        mw = self.mask_window
        single = _masks.gather(mw.mask_rects, np.arange(len(mw.mask_rects)))          # [36,4,4]
        preds_1 = self._predict(img, single)                                          # :70-72
        labels, counts = torch.from_numpy(preds_1.astype(np.int64)).unique(sorted=False, return_counts=True)
        label_majority = labels[counts.argmax()].item()                               # :74-75
        pred, preds_2 = label_majority, None
        if len(labels) == 1:                                                          # :78-79
            certifiable, preds_2 = self.robustness_certificate(img, pred)
        else:                                                                         # :80-90
            certifiable = False
            for label in labels.tolist():
                if label == label_majority:
                    continue
                for k in np.nonzero(preds_1 == label)[0]:
                    # second-round masking of the already one-masked image = double mask (k, j)
                    second = single.copy()
                    second[:, 2:4, :] = single[k, 0:2, :]
                    if (self._predict(img, second) == label).all():
                        pred = label
        if certify and preds_2 is None:                                               # :93-94
            preds_2 = self.robustness_certificate(img, label_majority)[1]
        return PatchCleanserRecord(pred, certifiable, preds_1.astype(np.int64),
                                   None if preds_2 is None else np.asarray(preds_2))

    def mask(self, img, msk):
        return img * msk + 0.5 * ~msk

    def robustness_certificate(self, img, label, batch_size=64):
        """All 630 double-mask predictions must equal `label` (:102-112).  `batch_size` is kept
        for signature compatibility; batching is the engine's `chunk`."""
        mw = self.mask_window
        preds = self._predict(img, _masks.gather(mw.double_mask_rects, np.arange(len(mw.double_mask_rects))))
        consistent = preds == label
        return bool(consistent.all()), consistent

    def reset(self):
        self.result = None

    def collect(self, records):
        self.result = PatchCleanserResult(records)


class PatchCleanserRecord(object):
    def __init__(self, pred, certifiable, preds_1, preds_2):
        self.prediction = pred
        self.certification = certifiable
        self.preds_1 = preds_1
        self.preds_2 = preds_2


class PatchCleanserResult(object):
    def __init__(self, records):
        self.predictions = np.stack([r.prediction for r in records])
        self.certifications = np.stack([r.certification for r in records])
        self.predictions_1 = np.stack([r.preds_1 for r in records])
        self.predictions_2 = [r.preds_2 for r in records]
