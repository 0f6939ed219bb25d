"""The classifier restatement (timm resnetv2_50x1_bit, not vendored in the reference) is
cross-checked against the HuggingFace port that ships in this image (transformers.models.bit)."""
import pytest
import torch

from oracle import resnetv2 as OR


def test_matches_hf_bit_port():
    tr = pytest.importorskip("transformers")
    from transformers import BitConfig, BitForImageClassification
    params = OR.random_init(seed=0, affine_jitter=0.1)
    hf = BitForImageClassification(BitConfig(num_labels=1000)).eval()
    sd = {"bit.embedder.convolution.weight": params["stem.conv.weight"], "bit.norm.weight": params["norm.weight"],
          "bit.norm.bias": params["norm.bias"], "classifier.1.weight": params["head.fc.weight"].reshape(1000, 2048),
          "classifier.1.bias": params["head.fc.bias"]}
    for k, v in params.items():
        if k.startswith("stages."):
            sd["bit.encoder." + k.replace(".blocks.", ".layers.")] = v
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    for H in (112, 224):
        z = torch.randn(2, 3, H, H, generator=torch.Generator().manual_seed(H))
        with torch.no_grad():
            a, b = OR.forward_normalized(params, z), hf(pixel_values=z).logits
        assert (a - b).abs().max().item() < 5e-5


def test_param_inventory():
    shapes = OR.param_shapes()
    assert len(shapes) == 153
    assert sum(int(torch.tensor(s).prod()) for s in shapes.values()) == 25549352   # 25.55 M parameters
    n_conv = sum(1 for k, s in shapes.items() if len(s) == 4 and k != "head.fc.weight")
    assert n_conv == 53
