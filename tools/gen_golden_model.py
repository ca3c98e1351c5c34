"""Generate tests/golden/model_*.npz from the UNMODIFIED reference (CPU, authoring container).

Weights: oracle.monodetr_torch.deterministic_state_dict() (keyed by parameter name, so the GPU box can rebuild
them without a checkpoint); inputs: synthetic_inputs(B, seed).  Outputs of the reference model are stored.
    python tools/gen_golden_model.py
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
warnings.filterwarnings("ignore")
import ref_shims  # noqa: E402
from oracle import monodetr_torch as om  # noqa: E402

KEYS = ("pred_logits", "pred_boxes", "pred_3d_dim", "pred_depth", "pred_angle", "pred_depth_map_logits")


def main():
    pkg = ref_shims.install()
    cfg = ref_shims.load_cfg()["model"]
    cfg["dropout"] = 0.0
    model, _ = pkg.build_monodetr(cfg)
    for m in model.modules():   # depth encoder hard-codes dropout=0.1 (depth_predictor.py:49-50): neutralise in memory
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    model.load_state_dict(om.with_aliases(om.deterministic_state_dict()))
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, B, H, W, training in (("model_eval_small", 1, 192, 640, False), ("model_eval_full", 1, 384, 1280, False),
                                    ("model_train_full", 1, 384, 1280, True)):
        model.train(training)
        images, calibs, sizes = om.synthetic_inputs(B, 0, H=H, W=W)
        with torch.no_grad():
            out = model(images, calibs, None, sizes)
        arrs = {k: out[k].numpy() for k in KEYS}
        for i, aux in enumerate(out["aux_outputs"]):
            for k, v in aux.items():
                arrs[f"aux{i}_{k}"] = v.numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), B=B, H=H, W=W, seed=0, training=int(training), **arrs)
        print(name, {k: v.shape for k, v in arrs.items() if not k.startswith("aux")})


if __name__ == "__main__":
    main()
