"""GPU micro-benchmark of the tcgen05 conv/linear family on the model's real shapes (B=8), both precision modes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monodetr_b200 import tc  # noqa: E402

B = 8
SHAPES = [  # name, H, W, Cin, Cout, k, stride
    ("l1.conv1 1x1 64->64", 96, 320, 64, 64, 1, 1),
    ("l1.conv2 3x3 64->64", 96, 320, 64, 64, 3, 1),
    ("l1.conv3 1x1 64->256", 96, 320, 64, 256, 1, 1),
    ("l1.conv1 1x1 256->64", 96, 320, 256, 64, 1, 1),
    ("l2.conv2 3x3 128->128", 48, 160, 128, 128, 3, 1),
    ("l2.conv3 1x1 128->512", 48, 160, 128, 512, 1, 1),
    ("l3.conv2 3x3 256->256", 24, 80, 256, 256, 3, 1),
    ("l3.conv3 1x1 256->1024", 24, 80, 256, 1024, 1, 1),
    ("l3.conv1 1x1 1024->256", 24, 80, 1024, 256, 1, 1),
    ("l4.conv2 3x3 512->512", 12, 40, 512, 512, 3, 1),
    ("enc linear 256->256 (M=81600)", 1, 10200, 256, 256, 1, 1),
]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us


for mode in tuple(os.environ["MDB_MODES"].split(",")) if os.environ.get("MDB_MODES") else (("tf32x3",) if os.environ.get("MDB_ONLY_X3") else ("bf16x3", "tf32x3", "tf32")):
    tc.set_precision(mode)
    print(f"== {mode}")
    for name, H, W, Cin, Cout, k, s in SHAPES:
        x = torch.randn(B, H, W, Cin, device="cuda")
        w = torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5
        wp = tc.split_weights([w])[0] if mode == "bf16x3" else tc.pack_weight(w)
        pad = k // 2
        y = tc.conv2d_forward(x, wp, None, None, k, k, s, pad)
        dy = torch.randn_like(y)
        res = torch.randn_like(y)
        flop = 2.0 * y.numel() * Cin * k * k
        byt = (x.numel() + y.numel()) * 4
        t_f = timeit(lambda: tc.conv2d_forward(x, wp, None, None, k, k, s, pad))
        t_fr = timeit(lambda: tc.conv2d_forward(x, wp, None, res, k, k, s, pad, relu=True))
        t_d = timeit(lambda: tc.conv2d_dgrad(dy, wp, x.shape, None, x, k, k, s, pad))
        t_w = timeit(lambda: tc.conv2d_wgrad(dy, x, None, k, k, s, pad))
        print(f"{name:34s} fwd {t_f:7.1f} us ({flop / t_f / 1e6:6.1f} TF/s, {byt / t_f / 1e3:6.0f} GB/s)  fwd+res {t_fr:7.1f}  "
              f"dgrad+mask {t_d:7.1f} ({flop / t_d / 1e6:6.1f} TF/s)  wgrad {t_w:7.1f} ({flop / t_w / 1e6:6.1f} TF/s)")
