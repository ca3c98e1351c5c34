"""Microbenchmark of the fused attention kernels at the model's three call shapes (B=8, 8 heads x 32):
depth-encoder self-attention (1920 x 1920), depth cross-attention (550 x 1920), group self-attention (11 groups of 50).
CUDA events, 20 iterations after 3 warm-ups; prints ms per call for forward and backward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from monodetr_b200 import kernels as K


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, B, Lq, Lk, drop, masked in [("depth-encoder self (train: dropout, no mask -- the model passes None)", 8, 1920, 1920, 0.1, False), ("depth cross (train)", 8, 550, 1920, 0.1, False), ("depth-encoder self, dropout + key_padding_mask", 8, 1920, 1920, 0.1, True),
                                          ("group self", 88, 50, 50, 0.1, False), ("depth-encoder self, eval", 8, 1920, 1920, 0.0, False)]:
        q = torch.randn(B, Lq, 256, device="cuda", generator=g)
        k = torch.randn(B, Lk, 256, device="cuda", generator=g)
        v = torch.randn(B, Lk, 256, device="cuda", generator=g)
        kpm = torch.zeros(B, Lk, dtype=torch.bool, device="cuda") if masked else None
        dout = torch.randn(B, Lq, 256, device="cuda", generator=g)
        o, lse, _ = K.attention_forward(q, k, v, kpm, drop_p=drop, site=1)
        tf = timeit(lambda: K.attention_forward(q, k, v, kpm, drop_p=drop, site=1))
        tb = timeit(lambda: K.attention_backward(q, k, v, kpm, o, lse, dout, drop_p=drop, site=1))
        flops = 4.0 * B * 8 * Lq * Lk * 32
        print(f"{name:72s} fwd {tf*1e3:8.1f} us ({flops/tf/1e9:6.1f} TF/s useful)   bwd {tb*1e3:8.1f} us ({2.5*flops/tb/1e9:6.1f} TF/s useful)")


if __name__ == "__main__":
    main()
