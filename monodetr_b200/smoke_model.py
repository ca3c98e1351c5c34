"""smoke(): one small forward+backward of the product model on cuda:0, forward checked against the CPU oracle."""
import torch


def run():
    from oracle import monodetr_torch as om            # checker only (allowed in smoke)
    from . import build_monodetr
    from .monodetr import DEFAULT_MODEL_CFG
    m, _ = build_monodetr(dict(DEFAULT_MODEL_CFG, dropout=0.0))
    sd = om.deterministic_state_dict()
    m.load_state_dict(om.with_aliases(sd))
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
    m = m.cuda().train()
    images, calibs, sizes = om.synthetic_inputs(1, 0, H=96, W=320)
    out = m(images.cuda(), calibs.cuda(), None, sizes.cuda())
    om.surrogate_loss(out).backward()
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = om.forward(sd, images, calibs, sizes, training=True)
    for k in ("pred_logits", "pred_boxes", "pred_3d_dim", "pred_depth", "pred_angle", "pred_depth_map_logits"):
        a, b = out[k].detach().float().cpu(), ref[k]
        err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
        assert err < 1e-3, (k, err)
    n = sum(1 for p in m.parameters() if p.grad is not None and torch.isfinite(p.grad).all())
    assert n == 313, n
