"""CPU: the HOST logic of the whole product model -- every Python wrapper and autograd function of monodetr_b200, with their
pointer / size / stride / flag arguments, the module wiring, the branch-stream fork / join code -- driven end to end on a
machine without a GPU through a stand-in for the device library at the C-ABI level (tests/fake_device_lib.py: each entry
point computes, on host memory, what include/monodetr_b200.h says the call computes), and compared with the oracle's
statement of the reference model (oracle/monodetr_torch.py, pinned to the unmodified reference by tests/test_oracle_model.py):
outputs and the gradient of every parameter.  The kernels themselves are pinned by the -m gpu suites against the same oracle.
"""
import pytest
import torch

from oracle import monodetr_torch as om
import fake_device_lib          # tests/fake_device_lib.py (pytest puts this directory on sys.path)


def _build(monkeypatch, precision="tf32x3"):
    fake = fake_device_lib.install(monkeypatch, {"tf32x3": 1, "bf16x3": 2}[precision])
    from monodetr_b200 import build_monodetr, tc
    from monodetr_b200.monodetr import DEFAULT_MODEL_CFG
    assert tc.get_precision() == precision
    m, _ = build_monodetr(dict(DEFAULT_MODEL_CFG, dropout=0.0, device="cpu"))
    sd = om.deterministic_state_dict()
    m.load_state_dict(om.with_aliases(sd))
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
    return fake, m, sd


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("precision,batch,deterministic", [("bf16x3", 1, False), ("tf32x3", 1, False), ("bf16x3", 2, True)])
def test_train_mode_forward_and_every_gradient_match_the_oracle(monkeypatch, precision, batch, deterministic):
    """bf16x3 = the default configuration: weights split once per forward (tc.prepacked / SplitW), the one-node encoder layer with
    its weight gradients forked onto the side stream; tf32x3: fp32 packed weights, the encoder layer as separate nodes; the
    reproducible mode (mdb_set_deterministic) sends every MSDeformAttn call down the two-step (pre-processing + op) path."""
    from monodetr_b200.bench_model import surrogate_loss
    fake, m, sd = _build(monkeypatch, precision)
    fake.deterministic = int(deterministic)
    from monodetr_b200 import functional as Fn
    node_calls, node = [], Fn.encoder_layer
    monkeypatch.setattr(Fn, "encoder_layer", lambda *a, **k: (node_calls.append(1), node(*a, **k))[1])
    m.train()
    images, calibs, sizes = om.synthetic_inputs(batch, 0, H=96, W=320)
    out = m(images, calibs, None, sizes)
    surrogate_loss(out).backward()

    sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}
    ref = om.forward(sdg, images, calibs, sizes, training=True)
    om.surrogate_loss(ref).backward()

    for k in ("pred_logits", "pred_boxes", "pred_3d_dim", "pred_depth", "pred_angle", "pred_depth_map_logits"):
        assert out[k].shape == ref[k].shape, k
        assert _rel(out[k].detach(), ref[k].detach()) < 1e-4, (k, _rel(out[k].detach(), ref[k].detach()))
    for a, b in zip(out["aux_outputs"], ref["aux_outputs"]):
        for k in a:
            assert _rel(a[k].detach(), b[k].detach()) < 1e-4, ("aux", k)

    by_name = om.with_aliases(sdg)          # (shared heads are listed under their decoder alias first)
    errs = []
    for name, p in m.named_parameters():
        want = by_name[name].grad
        if not p.requires_grad:                                     # frozen stem / layer1, depth_bin_values (depth_predictor.py:25)
            assert p.grad is None, name
            continue
        if p.grad is None:
            assert want is None or not want.any(), name            # the never-used tensors
            continue
        assert want is not None, name
        errs.append((_rel(p.grad, want), name, float(want.abs().max())))
    errs.sort()
    print("gradient errors (max-norm relative, per tensor): median %.2e; worst:" % errs[len(errs) // 2][0], errs[-8:])
    # Both sides are fp32 on the CPU with different operation orders (fused / summed projections, NHWC vs NCHW reductions): rounding
    # noise, amplified by ReLU selections and the piecewise-linear sampling exactly as on the GPU (tests/test_model_grad_gpu.py holds
    # the same 2e-2 class of bar); a wiring mistake shows up as an O(1) error on some tensor.  The key-projection biases of the
    # decoder's self-attention have an analytically ZERO gradient (softmax is invariant to a constant added to every key's score):
    # both sides produce ~1e-7 of noise there, which is compared absolutely.
    assert len(errs) == 313                                         # every gradient-receiving parameter of the model
    # bf16x3: the (hi, lo) weights are 2^-17 away from the oracle's fp32 weights, which flips a few more ReLU / max-pool selections
    # of this small (96 x 320, one image) problem: the same noise the GPU suite sees in that mode (median 3.3e-4 there).
    med_bar, worst_bar = (1e-3, 1e-1) if precision == "bf16x3" else (3e-4, 3e-2)
    assert errs[len(errs) // 2][0] < med_bar, errs[len(errs) // 2]
    for err, name, scale in errs:
        assert err < worst_bar or scale < 1e-6, (name, err, scale)
    assert len(node_calls) == (3 if precision == "bf16x3" and not deterministic else 0)      # the three encoder layers as one autograd node each
    if deterministic:
        assert fake.calls.get("mdb_msda_fused_forward_f32", 0) == 0 and fake.calls["mdb_msda_prep_forward_f32"] == 6
    # the wiring went through the library boundary, not around it
    conv = ("mdb_conv2d_forward_bf16x3", "mdb_conv2d_dgrad_bf16x3", "mdb_pack_gemm_weights_bf16x3") if precision == "bf16x3" else \
        ("mdb_conv2d_forward_f32", "mdb_conv2d_dgrad_f32", "mdb_pack_conv_weights_multi_f32")
    for fn in conv + ("mdb_conv2d_wgrad_bias_f32", "mdb_attention_forward_f32",
               "mdb_attention_backward_f32", "mdb_msda_prep_forward_f32",) + (() if deterministic else ("mdb_msda_fused_forward_f32", "mdb_msda_fused_backward_f32")) + (
               "mdb_add_layernorm_backward_f32", "mdb_groupnorm_backward_f32", "mdb_head_depth_backward_f32",
               "mdb_depth_tail_backward_f32", "mdb_box_refine_backward_f32", "mdb_stem_conv7x7_bn_relu_f32"):
        assert fake.calls.get(fn, 0) > 0, fn


def test_eval_mode_forward_matches_the_oracle(monkeypatch):
    _, m, sd = _build(monkeypatch)
    m.eval()
    images, calibs, sizes = om.synthetic_inputs(2, 1, H=96, W=320)
    with torch.no_grad():
        out = m(images, calibs, None, sizes)
        ref = om.forward(sd, images, calibs, sizes, training=False)
    for k in ("pred_logits", "pred_boxes", "pred_3d_dim", "pred_depth", "pred_angle", "pred_depth_map_logits"):
        assert out[k].shape == ref[k].shape, k
        assert _rel(out[k], ref[k]) < 1e-4, (k, _rel(out[k], ref[k]))
