"""One eager step of the model under `ncu --profile-from-start off` (cudaProfilerStart/Stop around step 2).
usage: profile_step.py [batch] [train|infer]   (infer = eval-mode forward only, BASELINE configs[4] uses batch 32)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monodetr_b200 import build_monodetr, kernels as K, tc  # noqa: E402
from monodetr_b200.bench_model import surrogate_loss, synthetic_batch  # noqa: E402
from monodetr_b200.monodetr import DEFAULT_MODEL_CFG  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tc.set_precision(os.environ.get("MDB_PRECISION", "bf16x3"))
torch.manual_seed(0)
model, _ = build_monodetr(DEFAULT_MODEL_CFG)
INFER = len(sys.argv) > 2 and sys.argv[2] == "infer"
model = model.cuda().eval() if INFER else model.cuda().train()
images, calibs, sizes = (t.cuda() for t in synthetic_batch(B, 1))


def step():
    if INFER:
        with torch.no_grad():
            surrogate_loss(model(images, calibs, None, sizes))
        return
    for p in model.parameters():
        p.grad = None
    out = model(images, calibs, None, sizes)
    surrogate_loss(out).backward()


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
