"""Run the attention core once per kernel (target for `ncu --set full -k regex:attn_ -c N`).
usage: one_attn.py [Lq Lk B drop_p]   (default: the depth encoder's 1920 x 1920 at B = 8 with dropout 0.1)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monodetr_b200 import kernels as K  # noqa: E402

Lq = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
Lk = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
drop = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(B, L, 256, device="cuda", generator=g) for L in (Lq, Lk, Lk))
dout = torch.randn(B, Lq, 256, device="cuda", generator=g)
for _ in range(2):
    o, lse, _ = K.attention_forward(q, k, v, None, drop_p=drop, site=1)
    K.attention_backward(q, k, v, None, o, lse, dout, drop_p=drop, site=1)
torch.cuda.synchronize()
