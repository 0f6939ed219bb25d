"""Process-wide registry of native engines (one per (device, image size, precision))."""
import os

import torch

_ENGINES = []


def register(engine):
    if engine not in _ENGINES:
        _ENGINES.append(engine)


def shared_engine(img, max_images=1):
    """An engine usable for weight-free helpers (paste, window sums): any live engine of the
    right geometry, else a minimal one."""
    dev = torch.cuda.current_device()
    for e in _ENGINES:
        if e.handle and e.img == img and e.device == dev and e.max_images >= max_images:
            return e
    from .engine import Engine
    e = Engine(img=img, precision=os.environ.get("DORPATCH_PRECISION", "tf32"), chunk=1,
               max_images=max(1, int(max_images)), autotune=False)
    _ENGINES.append(e)
    return e
