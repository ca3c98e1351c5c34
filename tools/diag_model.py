"""GPU diagnostic: stage-by-stage comparison of the product model against the CPU oracle (same weights/inputs)."""
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import monodetr_torch as om  # noqa: E402


def rel(a, b):
    a = a.detach().float().cpu(); b = b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def main():
    from monodetr_b200 import build_monodetr
    from monodetr_b200.monodetr import DEFAULT_MODEL_CFG
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (192, 640)
    training = len(sys.argv) > 3 and sys.argv[3] == "train"
    m, _ = build_monodetr(dict(DEFAULT_MODEL_CFG, dropout=0.0))
    sd = om.deterministic_state_dict()
    m.load_state_dict(om.with_aliases(sd))
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
    m = m.cuda().train(training)
    images, calibs, sizes = om.synthetic_inputs(1, 0, H=H, W=W)
    with torch.no_grad():
        t0 = time.time()
        feats_ref = om.backbone(sd, images)
        feats, pos = m.backbone(images.cuda())
        for i, (a, b) in enumerate(zip(feats, feats_ref)):
            print(f"backbone feat{i}: rel {rel(a.permute(0, 3, 1, 2), b):.3e}  shape {tuple(a.shape)}")
        srcs_ref = [om.conv_gn(sd, f"input_proj.{l}", f) for l, f in enumerate(feats_ref)]
        srcs_ref.append(om.conv_gn(sd, "input_proj.3", feats_ref[-1], stride=2, padding=1))
        srcs = [m.input_proj[l](f) for l, f in enumerate(feats)]
        srcs.append(m.input_proj[3](feats[-1]))
        for i, (a, b) in enumerate(zip(srcs, srcs_ref)):
            print(f"input_proj {i}: rel {rel(a.permute(0, 3, 1, 2), b):.3e}")
        # feed the ORACLE's srcs to the product modules from here on, so errors do not compound
        srcs_in = [s.permute(0, 2, 3, 1).contiguous().cuda() for s in srcs_ref]
        pos_ref = [om.position_embedding_sine(1, s.shape[2], s.shape[3], s.device) for s in srcs_ref]
        pos = [m.backbone[1](s) for s in srcs_in]
        for i, (a, b) in enumerate(zip(pos, pos_ref)):
            print(f"pos {i}: rel {rel(a, b[0].flatten(1).t()):.3e}")
        dl_r, dpe_r, wd_r, ip_r = om.depth_predictor(sd, srcs_ref, pos_ref[1])
        dl, dpe, wd, ip = m.depth_predictor(srcs_in, None, pos[1])
        print(f"depth logits rel {rel(dl.permute(0, 3, 1, 2), dl_r):.3e}; weighted depth rel {rel(wd, wd_r):.3e}; "
              f"depth_pos_embed rel {rel(dpe, dpe_r.flatten(2).transpose(1, 2)):.3e}; ip rel {rel(ip, ip_r.flatten(2).transpose(1, 2)):.3e}")
        nq = 550 if training else 50
        qe = sd["query_embed.weight"][:nq]
        hs_r, init_r, refs_r, dims_r = om.transformer(sd, srcs_ref, pos_ref, qe, dpe_r, training)
        hs, init, refs, dims, boxes = m.depthaware_transformer(srcs_in, None, pos, qe.cuda(), dpe_r.flatten(2).transpose(1, 2).contiguous().cuda(), None)
        print(f"hs rel {[round(rel(hs[i], hs_r[i]), 6) for i in range(3)]}; init_ref {rel(init, init_r):.3e}; refs {rel(refs, refs_r):.3e}; dims {rel(dims, dims_r):.3e}")
        out_r = om.forward(sd, images, calibs, sizes, training=training)
        out = m(images.cuda(), calibs.cuda(), None, sizes.cuda())
        for k in out_r:
            if k != "aux_outputs":
                print(f"END-TO-END {k}: rel {rel(out[k], out_r[k]):.3e}")
        print("time", time.time() - t0)


if __name__ == "__main__":
    main()
