"""Run one tcgen05 convolution shape a few times (target for `ncu --set full -k regex:tc_conv -s 3 -c 1`).
usage: one_conv.py Cin Cout k [fwd|fwdres|dgrad|wgrad] [H W]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monodetr_b200 import tc  # noqa: E402

Cin, Cout, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
what = sys.argv[4] if len(sys.argv) > 4 else "fwd"
H, W = (int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else (96, 320)
B = 8
x = torch.randn(B, H, W, Cin, device="cuda")
w = torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5
wp = tc.split_weights([w])[0] if tc.get_precision() == "bf16x3" else tc.pack_weight(w)
y = tc.conv2d_forward(x, wp, None, None, k, k, 1, k // 2)
dy, res = torch.randn_like(y), torch.randn_like(y)
for _ in range(5):
    if what == "fwd":
        tc.conv2d_forward(x, wp, None, None, k, k, 1, k // 2)
    elif what == "fwdres":
        tc.conv2d_forward(x, wp, None, res, k, k, 1, k // 2, relu=True)
    elif what == "dgrad":
        tc.conv2d_dgrad(dy, wp, x.shape, None, x, k, k, 1, k // 2)
    else:
        tc.conv2d_wgrad(dy, x, None, k, k, 1, k // 2)
torch.cuda.synchronize()
