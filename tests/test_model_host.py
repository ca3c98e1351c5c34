"""CPU: host-side contract of the product model (no compute): state_dict keys/shapes equal the reference's,
trainability rule, aliasing, and that the forward fails loudly without CUDA (no CPU fallback)."""
import os
import sys

import pytest
import torch

from monodetr_b200 import build_monodetr
from monodetr_b200.monodetr import DEFAULT_MODEL_CFG
from oracle import monodetr_torch as om

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(0)
    m, crit = build_monodetr(DEFAULT_MODEL_CFG)
    assert crit is None
    return m


def test_state_dict_contract(model):
    sd = model.state_dict()
    spec = om.with_aliases({k: torch.empty(s) for k, s in om.state_dict_spec().items()})
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in spec.items()}
    assert len(sd) == 582
    assert sum(p.numel() for p in model.parameters()) == 37675220
    trainable = [p for p in model.parameters() if p.requires_grad]
    assert len(trainable) == 328 and sum(p.numel() for p in trainable) == 37452739


def test_aliases_are_the_same_objects(model):
    assert model.depthaware_transformer.decoder.bbox_embed is model.bbox_embed
    assert model.depthaware_transformer.decoder.dim_embed is model.dim_embed_3d


def test_frozen_rule(model):
    for name, p in model.named_parameters():
        if name.startswith("backbone.0.body."):
            assert p.requires_grad == any(s in name for s in ("layer2", "layer3", "layer4")), name
    assert not model.depth_predictor.depth_bin_values.requires_grad


def test_init_rules(model):
    assert torch.allclose(model.class_embed[0].bias, torch.full((3,), -4.59511985))
    assert torch.all(model.bbox_embed[0].layers[-1].bias[2:] == -2.0)
    m = model.depthaware_transformer.encoder.layers[0].self_attn
    assert not m.sampling_offsets.weight.any() and not m.attention_weights.weight.any()
    assert float(m.sampling_offsets.bias.abs().max()) == 4.0


def test_strict_load_of_reference_shaped_checkpoint(model):
    sd = om.with_aliases(om.deterministic_state_dict())
    sd["backbone.0.body.bn1.num_batches_tracked"] = torch.tensor(0)           # dropped like the reference (backbone.py:41-50)
    model.load_state_dict(sd, strict=True)


def test_forward_requires_cuda(model):
    images, calibs, sizes = om.synthetic_inputs(1, 0, H=96, W=320)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(images, calibs, None, sizes)


@pytest.mark.reference
def test_state_dict_matches_unmodified_reference(model):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import warnings
    warnings.filterwarnings("ignore")
    import ref_shims
    pkg = ref_shims.install()
    ref, _ = pkg.build_monodetr(ref_shims.load_cfg()["model"])
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert a == b
    assert list(ref.state_dict().keys()) == list(model.state_dict().keys()) or set(a) == set(b)
    ref_train = {n for n, p in ref.named_parameters() if p.requires_grad}
    mine_train = {n for n, p in model.named_parameters() if p.requires_grad}
    assert ref_train == mine_train


@pytest.mark.reference
def test_build_returns_the_reference_criterion_when_importable():
    """B2: build_monodetr(cfg) -> (model, criterion) like monodetr.py:550-614; with the reference package importable
    (as inside tools/train_val.py) the criterion is the reference's SetCriterion with the reference's weight_dict."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import warnings
    warnings.filterwarnings("ignore")
    import ref_shims
    pkg = ref_shims.install()
    cfg = ref_shims.load_cfg()["model"]
    _, ref_crit = pkg.build_monodetr(cfg)
    from monodetr_b200 import build_monodetr
    _, crit = build_monodetr(cfg)
    assert type(crit) is type(ref_crit)
    assert crit.weight_dict == ref_crit.weight_dict and crit.losses == ref_crit.losses
    assert crit.focal_alpha == ref_crit.focal_alpha and crit.num_classes == ref_crit.num_classes
    # without loss weights in the cfg (stand-alone model use) there is nothing to build a criterion from
    from monodetr_b200.monodetr import DEFAULT_MODEL_CFG
    assert build_monodetr(DEFAULT_MODEL_CFG)[1] is None
