"""Live pin: run the UNMODIFIED reference (/root/reference, CPU shim) next to the oracle.
Only possible in the build container -- skipped wherever the reference is absent (the GPU box),
where tests/test_oracle_golden.py checks the same things against the frozen vectors."""
import contextlib
import importlib.util
import io
import os
import random

import numpy as np
import pytest
import torch

from oracle import attack as OA
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
HERE = os.path.dirname(os.path.abspath(__file__))


def _mod():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_generate_live_bit_exact(tmp_path, monkeypatch):
    mg = _mod()
    tiny = mg.TinyNet().eval()
    xs = torch.rand(1, 3, 56, 56, generator=torch.Generator().manual_seed(21))
    kw = dict(patch_budget=0.12, n_classes=1000, targeted=True, y=torch.tensor([7]), max_iterations=25, sampling_size=3,
              dropout=2)
    with ref_shim.reference_modules() as ref:
        m_ref, p_ref, log_ref, rn, rt = mg.run_generate(ref, tiny, xs, **kw)
    random.seed(1234); torch.manual_seed(1234); np.random.seed(1234)
    log = []
    m, p = OA.generate(tiny, xs, log=log.append, **kw)
    assert np.array_equal(m.numpy(), m_ref) and np.array_equal(p.numpy(), p_ref)
    assert log == log_ref
    assert np.array_equal(np.random.get_state()[1][:4], rn)


def test_golden_file_is_reproducible():
    """The committed fixture equals what the reference produces now (spot check: G2 + G4)."""
    G = np.load(os.path.join(HERE, "golden", "reference_golden.npz"))
    with ref_shim.reference_modules() as ref:
        x, m, p = (torch.from_numpy(G[k]) for k in ("g2_x", "g2_m", "g2_p"))
        d = ref.utils.clip(m, p, x, 4.0)
        assert np.array_equal(d.numpy(), G["g2_delta"])
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            sel = ref.attack.DorPatch().patch_selection(torch.from_numpy(G["g4_mask"]), 0.10)
        assert np.array_equal(np.packbits(sel.numpy().astype(bool)), G["g4_sel_0.1"])
