"""Golden vectors for the inference post-process, produced by the UNMODIFIED reference functions
(lib/helpers/decode_helper.py: extract_dets_from_outputs, decode_detections; lib/datasets/kitti/kitti_utils.py: Calibration).
Run in the build container (needs /root/reference):   python tools/gen_golden_decode.py   -> tests/golden/decode.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, "/root/reference")
from oracle.decode import synthetic_heads  # noqa: E402

from lib.helpers.decode_helper import decode_detections, extract_dets_from_outputs  # noqa: E402
from lib.datasets.kitti.kitti_utils import Calibration  # noqa: E402

out = {}
for name, (seed, B, Q, topk, thr) in {"eval": (11, 4, 50, 50, 0.2), "train_queries": (12, 2, 550, 50, 0.2),
                                      "low_threshold": (13, 3, 50, 20, 0.0)}.items():
    h = synthetic_heads(seed, B, Q)
    outputs = {"pred_logits": torch.from_numpy(h["logits"]), "pred_boxes": torch.from_numpy(h["boxes"]),
               "pred_angle": torch.from_numpy(h["angle"]), "pred_3d_dim": torch.from_numpy(h["dim3"]),
               "pred_depth": torch.from_numpy(h["depth"])}
    dets = extract_dets_from_outputs(outputs, K=50, topk=topk).numpy()
    calibs = [Calibration({"P2": h["P2"][i], "R0": np.eye(3, dtype=np.float32), "Tr_velo2cam": np.zeros((3, 4), np.float32)})
              for i in range(B)]
    info = {"img_id": np.arange(B), "img_size": h["img_size"]}
    res = decode_detections(dets.copy(), info, calibs, h["mean_size"].astype(np.float64), thr)
    rows = np.zeros((B, topk, 14), np.float64)
    count = np.zeros(B, np.int64)
    for i in range(B):
        count[i] = len(res[i])
        if res[i]:
            rows[i, :len(res[i])] = np.array(res[i], np.float64)
    out[f"{name}.cfg"] = np.array([seed, B, Q, topk, thr], np.float64)
    out[f"{name}.dets"] = dets
    out[f"{name}.rows"] = rows
    out[f"{name}.count"] = count
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "decode.npz"), **out)
print({k: v.shape for k, v in out.items()})
