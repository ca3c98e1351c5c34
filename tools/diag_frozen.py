"""Gradient parity diagnostics: sm_100a path (both tensor-core modes) and the fp32 CPU oracle, each against the fp64 CPU oracle;
max-norm and L2 relative error per tensor, worst per stage.  usage: diag_frozen.py [seed] [freeze 0|1]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import monodetr_torch as om  # noqa: E402
from monodetr_b200 import build_monodetr, tc  # noqa: E402
from monodetr_b200.monodetr import DEFAULT_MODEL_CFG  # noqa: E402
from monodetr_b200.ms_deform_attn import MSDeformAttn  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 11
freeze = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
images, calibs, sizes = om.synthetic_inputs(1, seed, H=192, W=640)
STAGES = ("backbone.0.body.layer2", "backbone.0.body.layer3", "backbone.0.body.layer4", "input_proj", "depth_predictor",
          "depthaware_transformer.encoder", "depthaware_transformer.decoder", "depthaware_transformer", "query_embed", "class_embed",
          "bbox_embed", "dim_embed_3d", "angle_embed", "depth_embed")


def ours(mode):
    tc.set_precision(mode)
    m, _ = build_monodetr(dict(DEFAULT_MODEL_CFG, dropout=0.0))
    m.load_state_dict(om.with_aliases(om.deterministic_state_dict()))
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
    m = m.cuda().train()
    MSDeformAttn.freeze_sampling_locations = freeze
    out = m(images.cuda(), calibs.cuda(), None, sizes.cuda())
    om.surrogate_loss(out).backward()
    torch.cuda.synchronize()
    MSDeformAttn.freeze_sampling_locations = False
    return {n: p.grad.double().cpu() for n, p in m.named_parameters() if p.grad is not None}, {k: v.detach().double().cpu() for k, v in out.items() if torch.is_tensor(v)}


def ref(dt):
    om.FREEZE_SAMPLING = freeze
    sd = {k: (v.to(dt) if v.dtype.is_floating_point else v.clone()).requires_grad_(v.dtype.is_floating_point)
          for k, v in om.deterministic_state_dict().items()}
    out = om.forward(sd, images.to(dt), calibs.to(dt), sizes.to(dt), training=True)
    om.surrogate_loss(out).backward()
    om.FREEZE_SAMPLING = False
    return {k: v.grad.double() for k, v in sd.items() if v.grad is not None}, {k: v.detach().double() for k, v in out.items() if torch.is_tensor(v)}


def cmp(a, b, tag):
    rows = []
    for k in a:
        if k not in b or "decoder.bbox_embed" in k or "decoder.dim_embed" in k or "running" in k:
            continue
        s = float(b[k].abs().max())
        if s < 1e-7:
            continue
        d = a[k] - b[k]
        rows.append((float(d.abs().max()) / s, float(d.norm() / b[k].norm()), k))
    mx = np.array([r[0] for r in rows]); l2 = np.array([r[1] for r in rows])
    print(f"{tag}: n {len(rows)}  max-norm median {np.median(mx):.2e} worst {mx.max():.2e} | L2 median {np.median(l2):.2e} worst {l2.max():.2e}")
    per = {}
    for m_, l_, k in rows:
        st = next((s for s in STAGES if k.startswith(s)), "other")
        cur = per.get(st, (0, 0))
        per[st] = (max(cur[0], m_), max(cur[1], l_))
    print("   per stage (max-norm, L2):", {k: "%.1e %.1e" % v for k, v in per.items()})


g64, o64 = ref(torch.float64)
g32, o32 = ref(torch.float32)
print("outputs fp32 oracle vs fp64:", {k: "%.1e" % float((o32[k] - o64[k]).abs().max()) for k in o64})
cmp(g32, g64, "fp32 oracle vs fp64 oracle")
for mode in ("bf16x3", "tf32x3"):
    g, o = ours(mode)
    print(f"outputs ours[{mode}] vs fp64:", {k: "%.1e" % float((o[k] - o64[k]).abs().max()) for k in o64})
    cmp(g, g64, f"ours[{mode}] vs fp64 oracle")
os.environ["MDB_ATTN_LEGACY"] = "1"
