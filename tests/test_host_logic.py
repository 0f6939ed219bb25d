"""CPU tests of the product's host-side logic (no GPU, no compute calls into the library):
mask geometry, the scalar state machine, target selection, result paths, CLI surface, the C-ABI
export list, and "no CPU fallback" behaviour."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import attack as OA
from oracle import masks as OM
from oracle import resnetv2 as OR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mask_tables_equal_oracle():
    from dorpatch_b200 import masks as PM
    for H in (56, 112, 224):
        for d in (1, 2):
            assert np.array_equal(PM.universe(H, d), OM.rects_to_array(OM.universe_rects(H, d)))
    t = PM.universe(56, 2)
    assert np.array_equal(PM.to_bool(t[:50], 56), OM.rects_to_bool(OM.universe_rects(56, 2)[:50], 56))
    g = PM.gather(t, np.array([[0, 5]]), np.array([[7, 9]]))
    assert g.shape == (1, 2, 4, 4) and np.array_equal(g[0, 1, 2:], t[9])


def test_pick_target_equals_oracle_set_target():
    from dorpatch_b200.attack import _pick_target
    rng = np.random.RandomState(0)
    for _ in range(300):
        n = rng.randint(1, 9)
        preds = rng.randint(0, 5, n)
        label = int(rng.randint(0, 5))
        assert _pick_target(preds, label) == OA.set_target(torch.from_numpy(preds), label)


def test_bookkeeping_state_machine_equals_oracle():
    """Random loss sequences through both state machines: identical lr / coefficients /
    failed sets / decisions at every step (patience, lr decay, early stop included)."""
    from dorpatch_b200.attack import _ImageState
    for seed in range(4):
        rng = np.random.RandomState(seed)
        a = _ImageState(0.01, 1e-3, 3, False, np.random.RandomState(seed))
        b = OA.ImageState(0.01, 1e-3, 3, False, np.random.RandomState(seed))
        for stage in (0, 1):
            a.reset()
            b.reset_stage()
            for i in range(1300):
                ia, na = a.sample(i, 144, 8)
                ib, nb = OA.sample_indices(b, i, 144, 8)
                assert na == nb and np.array_equal(ia, ib)
                loss = (rng.rand(8) < (0.3 if i < 900 else 0.9)).astype(np.float32) * 0.5 * (i < 1100 or rng.rand() < 0.5)
                tgt = np.float32(10.0 / (1 + 0.01 * (i % 450)) + rng.rand() * 1e-3)
                ra = a.bookkeeping(stage, i, loss, ia, na, tgt)
                rb = OA.bookkeeping(b, stage, i, loss, ib, nb, tgt)
                assert ra == rb
                assert (a.lr, a.structured, a.coeff_group_lasso, a.failed, a.not_decay, a.num_failure) == \
                       (b.lr, b.structured, b.coeff_group_lasso, b.failed, b.not_decay, b.num_failure)
                if ra[1]:
                    break


def test_fp32_lr_decay_needs_two_decays():
    """0.01f * 0.1f == float32(1e-3), which is NOT < 1e-3: the earliest stop is after the second decay."""
    from dorpatch_b200.attack import _ImageState
    s = _ImageState(0.01, 1e-3, 0, True, np.random.RandomState(0))
    stops = []
    for i in range(405):
        _, stop = s.bookkeeping(1, i, np.zeros(2, np.float32), np.array([0, 1]), 0, 1.0)
        stops.append(stop)
    assert stops.index(True) == 402


def test_c_abi_exports_every_declared_symbol():
    from dorpatch_b200 import _lib
    header = open(os.path.join(ROOT, "include", "dorpatch.h")).read()
    declared = set(re.findall(r"\b(dp_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = _lib.load()
    assert lib.dp_abi_version() == int(re.search(r"#define DP_ABI_VERSION (\d+)", header).group(1))
    for name in declared:
        assert hasattr(lib, name), "libdorpatch.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_no_cpu_fallback():
    from dorpatch_b200.attack import DorPatch
    from dorpatch_b200.resnetv2 import ResNetV2
    from dorpatch_b200.utils import NormModel, get_normalize
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    net = ResNetV2()
    model = torch.nn.DataParallel(NormModel(net, get_normalize("imagenet", "resnetv2")))
    x = torch.rand(1, 3, 56, 56)
    with pytest.raises(RuntimeError):
        DorPatch().generate(model, x, 0.1, 1000, "results/a/b", 0)
    with pytest.raises(RuntimeError):
        net(x)
    with pytest.raises(TypeError):          # an unsupported classifier is rejected, not silently run in eager mode
        DorPatch().generate(torch.nn.Linear(3, 3), x, 0.1, 1000, "results/a/b", 0)
    from dorpatch_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(img=56)


def test_model_container_matches_oracle_params():
    from dorpatch_b200.resnetv2 import ResNetV2
    sd = ResNetV2(seed=0).state_dict()
    ref = OR.random_init(seed=0)
    assert list(sd.keys()) == list(ref.keys())
    assert all(torch.equal(sd[k], ref[k]) for k in ref)


def test_result_path_and_cli_surface(tmp_path, monkeypatch):
    from dorpatch_b200 import main as M
    from dorpatch_b200.utils import generate_saving_path
    monkeypatch.chdir(tmp_path)
    args = M.parser.parse_args(["--targeted", "--dropout", "1", "--synthetic", "3", "--max_iterations", "7"])
    assert (args.batch_size, args.epsilon, args.lr, args.patch_budget, args.num_patch) == (1, 4.0, 0.01, 0.12, -1)
    p = generate_saving_path(vars(args).copy())
    ref_top = "results/dataset=imagenet_base_arch=resnetv2_targeted=True_attack=DorPatch_dropout=1_density=0.001_structured=0.001"
    # flags that change the artefacts (synthetic data, truncated run) may not share the real run's directory
    assert p == ref_top + "__synthetic=3_max_iterations=7/num_patch=-1_patch_budget=0.12"
    assert os.path.isdir(p)
    # with the reference's own flag values the path is the reference's (utils.py:24-44), whatever precision / chunk say
    args = M.parser.parse_args(["--targeted", "--dropout", "1", "--precision", "bf16", "--chunk", "64", "--num_batches", "3"])
    assert generate_saving_path(vars(args).copy()) == ref_top + "/num_patch=-1_patch_budget=0.12"
    ref_flags = ["--device", "--dataset", "--data_dir", "--model_dir", "--base_arch", "--targeted", "--patch_budget",
                 "--attack", "--batch-size", "--epsilon", "--lr", "--num_patch", "--dropout", "--density", "--structured"]
    have = {o for a in M.parser._actions for o in a.option_strings}
    assert all(f in have for f in ref_flags)


def test_reference_module_names_resolve():
    import attack
    import utils
    from defenses.PatchCleanser import MaskWindow, PatchCleanser, PatchCleanserRecord, PatchCleanserResult  # noqa: F401
    assert hasattr(attack, "DorPatch") and hasattr(attack.DorPatch, "generate")
    for n in ("clip", "NormModel", "get_model", "get_dataset", "set_random_seed", "set_device", "generate_saving_path",
              "NUM_CLASSES_DICT", "convert_float_list_to_str"):
        assert hasattr(utils, n), n
    import pickle
    rec = PatchCleanserRecord(3, True, np.arange(36), np.ones(630, bool))
    back = pickle.loads(pickle.dumps([[rec]]))
    assert back[0][0].prediction == 3 and back[0][0].preds_2.shape == (630,)


def test_bool_universe_round_trips_to_rectangles():
    """collect_failure's reference signature takes the bool mask universe (attack.py:384-406); the native path needs
    rectangles: masks.from_bool must invert masks.to_bool on the whole universe, and reject anything else."""
    from dorpatch_b200 import masks as PM
    for dropout in (1, 2):
        table = PM.universe(112, dropout)
        dense = PM.to_bool(table, 112)
        sel = np.arange(len(table)) if dropout == 1 else np.random.RandomState(0).choice(len(table), 300, replace=False)
        back = PM.from_bool(dense[sel])
        assert np.array_equal(PM.to_bool(back, 112), dense[sel])
    bad = np.ones((1, 1, 112, 112), bool)
    bad[0, 0, 3:9, 3:9] = False
    bad[0, 0, 20:30, 40:50] = False
    bad[0, 0, 60:70, 5:15] = False                # three rectangles: no dense-mask path exists
    with pytest.raises(NotImplementedError):
        PM.from_bool(bad)


def test_sampling_prefetch_never_changes_rng_consumption():
    """DorPatch.generate draws step i+1's samples on a helper thread while the GPU runs step i whenever _may_prefetch says
    the reference's RNG consumption cannot change.  Simulation of the loop's control flow (sampling, bookkeeping with lr
    decays and early stops, stage restart) with and without prefetching over adversarial loss sequences: identical index
    sequences, identical set of (image, step) draws, identical final RNG states."""
    from dorpatch_b200.attack import _ImageState, _may_prefetch, PATIENCE

    def run(prefetch, seed, B=3, S=6, n_mask=144, max_iterations=700):
        rngs = [np.random.RandomState(100 + b) for b in range(B)]
        st = [_ImageState(0.01, 1e-3, 1, False, rngs[b]) for b in range(B)]
        lossgen = np.random.RandomState(seed)
        log = []

        def draw(i):
            out = []
            for b, s in enumerate(st):
                if s.active:
                    out.append((b, i, tuple(s.sample(i, n_mask, S)[0])))
            return out
        for stage in range(2):
            for s in st:
                s.reset()
            pending = None
            for i in range(max_iterations):
                if pending is not None and pending[0] == i:
                    got = pending[1]
                else:
                    got = draw(i)
                pending = None
                if prefetch and _may_prefetch(st, i, max_iterations):
                    pending = (i + 1, draw(i + 1))          # drawn BEFORE this step's bookkeeping, as the helper thread may
                log.extend(got)
                # loss_target: image 0 improves for a while then plateaus (decays + stop), image 1 improves slowly, image 2 is noisy
                for b, s in enumerate(st):
                    if not s.active:
                        continue
                    target = {0: max(5.0 - 0.05 * i, 2.0), 1: 10.0 - 0.002 * i, 2: 3.0 + lossgen.rand()}[b]
                    loss = (lossgen.rand(S) < 0.5).astype(np.float32)
                    improved, stop = s.bookkeeping(stage, i, loss, np.arange(S), 0, target)
                    if stop:
                        s.active = False
                if not any(s.active for s in st):
                    break
        return log, [r.get_state()[1].tolist() + [r.get_state()[2]] for r in rngs], [s.active for s in st]

    for seed in (0, 1, 2):
        a = run(False, seed)
        b = run(True, seed)
        assert a[0] == b[0]                                   # same draws, same order, same values
        assert a[1] == b[1]                                   # same final RNG states: nothing extra was consumed
        assert a[2] == b[2]
    # the scenario does exercise stops and plateaus (otherwise the test proves nothing)
    log, _, active = run(True, 0)
    assert not all(active) or len({(b, i) for b, i, _ in log}) < 2 * 3 * 700


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm: the oracle port of the reference's step on the host cores) must
    print ONE JSON line carrying the contract's keys, with metric / unit / config equal to the native arm's."""
    import json
    import subprocess
    import sys
    root = ROOT
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "EOT-samples/sec" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "configs[2]" in d["config"]["workload"]
