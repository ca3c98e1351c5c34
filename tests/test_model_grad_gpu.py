"""Gradient parity of the whole model, with and without the discontinuity of the sampling op, and dropout statistics.

1. FROZEN SAMPLING LOCATIONS.  d(bilinear)/d(location) jumps at cell borders, so gradients that flow THROUGH the sampling
   locations could amplify forward noise.  Both sides can treat the locations as constants in backward (CPU oracle:
   `oracle.monodetr_torch.FREEZE_SAMPLING`, sm_100a path: `MSDeformAttn.freeze_sampling_locations`); every remaining gradient
   -- backbone, neck, depth predictor, encoder, decoder, heads, incl. the gradient wrt value and attention weights of every
   MSDeformAttn -- is then compared per stage, at 192x640 (B=1) and at the benchmark's 1280x384 (B=2), and the unfrozen
   comparison runs beside it.  MEASURED (tools/diag_frozen.py, profiles/r02_gradient_parity.txt): freezing changes nothing --
   median relative error 3.3e-4 either way, worst tensors 5e-3..1e-2 in the BACKBONE, not in the sampling projections.  The
   tail comes from ReLU / max-pool selections flipping on 1e-6 forward noise: PyTorch's own fp32 CPU path against an fp64
   run of the same oracle shows the same tail (worst 5.5e-3, backbone layer3) with a median of 9e-6.  The median of the
   sm_100a path (3e-4) is the tensor cores' truncating fp32 accumulation compounding over ~50 layers; single-pass TF32, the
   reference's own default on this hardware, sits at 1e-2.  Bars: median < 1e-3 (L2 and max-norm), every tensor < 2e-2
   max-norm and L2 -- i.e. all non-MSDA gradients hold 1e-3 at the median and the tail stays within 2x of what a pure
   fp32 implementation shows against fp64.
2. DROPOUT ON.  Hash masks (ours) and torch's RNG (oracle, `DROPOUT_P`) cannot agree element-wise; the ensemble statistics
   of the outputs over seeds must: same keep-probability at the same 34 sites -> same output distribution.
"""
import numpy as np
import pytest
import torch

from oracle import monodetr_torch as om

pytestmark = pytest.mark.gpu
STAGES = ("backbone", "input_proj", "depth_predictor", "depthaware_transformer.encoder", "depthaware_transformer.decoder",
          "depthaware_transformer.level_embed", "depthaware_transformer.reference_points", "query_embed", "class_embed",
          "bbox_embed", "dim_embed_3d", "angle_embed", "depth_embed")


def _model(dropout=0.0):
    from monodetr_b200 import build_monodetr
    from monodetr_b200.monodetr import DEFAULT_MODEL_CFG
    m, _ = build_monodetr(dict(DEFAULT_MODEL_CFG, dropout=dropout))
    m.load_state_dict(om.with_aliases(om.deterministic_state_dict()))
    if dropout == 0.0:                  # sites whose rate the cfg does not reach (the depth encoder's fixed 0.1)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, torch.nn.MultiheadAttention):
                mod.dropout = 0.0
    return m.cuda()


def _grad_report(m, sd):
    per_stage, rel_max, rel_l2 = {}, [], []
    for name, p in m.named_parameters():
        if name.startswith("depthaware_transformer.decoder.bbox_embed") or name.startswith("depthaware_transformer.decoder.dim_embed"):
            continue                                            # aliases of bbox_embed / dim_embed_3d
        if not p.requires_grad or p.grad is None:
            continue
        gref = sd[name].grad
        assert gref is not None, name
        scale = float(gref.abs().max())
        if scale < 1e-7:
            continue
        d = p.grad.cpu() - gref
        r, l2 = float(d.abs().max()) / scale, float(d.norm() / gref.norm())
        rel_max.append(r)
        rel_l2.append(l2)
        stage = next(s for s in STAGES if name.startswith(s))
        cur = per_stage.get(stage, (0.0, 0.0, ""))
        per_stage[stage] = (max(cur[0], r), max(cur[1], l2), name if r > cur[0] else cur[2])
    return per_stage, rel_max, rel_l2


@pytest.mark.parametrize("B,H,W,freeze", [(1, 192, 640, True), (1, 192, 640, False), (2, 384, 1280, True)])
def test_gradients_per_stage(B, H, W, freeze):
    from monodetr_b200.ms_deform_attn import MSDeformAttn
    m = _model(0.0).train()
    images, calibs, sizes = om.synthetic_inputs(B, 11, H=H, W=W)
    MSDeformAttn.freeze_sampling_locations = freeze
    om.FREEZE_SAMPLING = freeze
    try:
        out = m(images.cuda(), calibs.cuda(), None, sizes.cuda())
        om.surrogate_loss(out).backward()
        torch.cuda.synchronize()
        sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in om.deterministic_state_dict().items()}
        ref_out = om.forward(sd, images, calibs, sizes, training=True)
        om.surrogate_loss(ref_out).backward()
    finally:
        MSDeformAttn.freeze_sampling_locations = False
        om.FREEZE_SAMPLING = False
    per_stage, rel_max, rel_l2 = _grad_report(m, sd)
    print(f"B={B} {H}x{W} frozen={freeze}: worst relative gradient error per stage (max-norm, L2):",
          {k: f"{v[0]:.1e} {v[1]:.1e} ({v[2].split('.')[-2]})" for k, v in per_stage.items()},
          "median", f"{float(np.median(rel_max)):.2e} {float(np.median(rel_l2)):.2e}", "tensors", len(rel_max))
    assert len(rel_max) > (240 if freeze else 270)              # (frozen: the sampling_offsets projections get no gradient)
    assert float(np.median(rel_max)) < 1e-3 and float(np.median(rel_l2)) < 1e-3
    for stage, (r, l2, name) in per_stage.items():
        assert r < 2e-2 and l2 < 2e-2, (stage, name, r, l2)


def test_dropout_output_statistics_match_the_oracle():
    from monodetr_b200 import kernels as K
    N, p = 32, 0.1
    images, calibs, sizes = om.synthetic_inputs(2, 5, H=96, W=320)
    m = _model(p).train()
    dev = torch.device("cuda", torch.cuda.current_device())
    keys = ("pred_logits", "pred_boxes", "pred_3d_dim", "pred_depth", "pred_angle")

    def stats(out):
        return [float(out[k].float().mean()) for k in keys] + [float(out[k].float().std()) for k in keys]

    ours, ref = [], []
    with torch.no_grad():
        for s in range(N):
            K.reseed(dev, 1000 + s)
            ours.append(stats(m(images.cuda(), calibs.cuda(), None, sizes.cuda())))
        sd = om.deterministic_state_dict()
        om.DROPOUT_P = p
        try:
            for s in range(N):
                torch.manual_seed(2000 + s)
                ref.append(stats(om.forward(sd, images, calibs, sizes, training=True)))
        finally:
            om.DROPOUT_P = 0.0
    ours, ref = np.array(ours), np.array(ref)
    assert ours.std(0).min() > 0                                # masks really differ from seed to seed
    se = np.sqrt(ours.var(0, ddof=1) / N + ref.var(0, ddof=1) / N)
    z = np.abs(ours.mean(0) - ref.mean(0)) / (se + 1e-6 * np.abs(ref.mean(0)) + 1e-9)
    print("dropout ensemble z-scores (mean / std of each head output):", np.round(z, 2))
    assert z.max() < 5.0, z
