"""One call of each "next row" (criterion fwd+bwd, post-process, pre-process) between cudaProfilerStart/Stop, for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python tools/profile_next_rows.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import criterion as oc, decode as od, preprocess as op  # noqa: E402
from monodetr_b200.criterion import build_criterion  # noqa: E402
from monodetr_b200.decode import decode_detections_device, extract_dets_from_outputs  # noqa: E402
from monodetr_b200.preprocess import ImageBatchPreprocessor, get_affine_transform  # noqa: E402

dev = torch.device("cuda")
cfg = {"num_classes": 3, "cls_loss_coef": 2, "focal_alpha": 0.25, "bbox_loss_coef": 5, "giou_loss_coef": 2, "3dcenter_loss_coef": 10,
       "dim_loss_coef": 1, "angle_loss_coef": 1, "depth_loss_coef": 1, "depth_map_loss_coef": 1, "set_cost_class": 2, "set_cost_bbox": 5,
       "set_cost_giou": 2, "set_cost_3dcenter": 10, "aux_loss": True, "dec_layers": 3}
out, padded = oc.synthetic_case(5, 8, 550)
crit = build_criterion(cfg).to(dev).train()
o = {k: (v.to(dev).requires_grad_(True) if torch.is_tensor(v) else [{kk: vv.to(dev).requires_grad_(True) for kk, vv in a.items()} for a in v])
     for k, v in out.items()}
tg = {k: v.to(dev) for k, v in padded.items()}
h = od.synthetic_heads(3, 32, 50)
ho = {"pred_logits": torch.from_numpy(h["logits"]).to(dev), "pred_boxes": torch.from_numpy(h["boxes"]).to(dev),
      "pred_3d_dim": torch.from_numpy(h["dim3"]).to(dev), "pred_depth": torch.from_numpy(h["depth"]).to(dev),
      "pred_angle": torch.from_numpy(h["angle"]).to(dev)}
sz, P2, ms3 = torch.from_numpy(h["img_size"]).to(dev), torch.from_numpy(h["P2"]).to(dev), torch.from_numpy(h["mean_size"]).to(dev)
sizes = [(1242, 375), (1224, 370), (1238, 374), (1241, 376)] * 2
imgs = op.synthetic_images(1, sizes)
tinv = np.stack([get_affine_transform(np.array(s, np.float64) / 2, np.array(s, np.float64), 0, np.array([1280, 384]), inv=1)[1] for s in sizes])
src = [torch.from_numpy(im).to(dev) for im in imgs]
pre = ImageBatchPreprocessor((1280, 384), device=dev)


def once():
    crit(o, tg)
    crit.weighted_sum().backward()
    decode_detections_device(extract_dets_from_outputs(ho, topk=50), sz, P2, ms3, 0.2)
    pre(src, tinv)


once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
