"""Time the MSDeformAttn backward in the default (vector reductions) and the reproducible (ordered accumulation) mode at the
encoder shape (B = 8, Lq = 10 200); prints one line."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import monodetr_b200                                                       # noqa: E402
from bench import make_msda_inputs                                         # noqa: E402
from monodetr_b200.msda import ms_deform_attn_backward                     # noqa: E402

dv = make_msda_inputs(8, 10200, seed=0, device="cuda")


def ms(n):
    for _ in range(2):
        ms_deform_attn_backward(*dv, 64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ms_deform_attn_backward(*dv, 64)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


default = ms(10)
monodetr_b200.set_deterministic(True)
ordered = ms(3)
print(f"msda backward B=8 Lq=10200: default {default:.3f} ms, reproducible mode {ordered:.3f} ms ({ordered / default:.1f}x)")
