"""Generate tests/golden/msda_*.npz from the UNMODIFIED reference (CPU).

Run in the authoring container (needs /root/reference):  python tools/gen_golden_msda.py
Each fixture holds seeded inputs and the outputs/gradients of the reference's own
ms_deform_attn_core_pytorch (lib/models/monodetr/ops/functions/ms_deform_attn_func.py:41-61)
differentiated by autograd.  Case `reftest_*` reproduces the generator of the reference's
ops/test.py:21-37 (seed 3, shapes [(6,4),(3,2)], N=1 M=2 D=2 Lq=L=P=2).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(__file__))
import ref_shims  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_case(seed, shapes, N, M, D, Lq, P, dtype, loc_lo=0.0, loc_hi=1.0, value_scale=1.0, reftest=False):
    torch.manual_seed(seed)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    S = int(shapes_t.prod(1).sum())
    L = len(shapes)
    if reftest:  # ops/test.py:33-36
        value = torch.rand(N, S, M, D) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        attn = torch.rand(N, Lq, M, L, P) + 1e-5
        attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    else:
        value = torch.randn(N, S, M, D) * value_scale
        loc = torch.rand(N, Lq, M, L, P, 2) * (loc_hi - loc_lo) + loc_lo
        attn = torch.softmax(torch.randn(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
    grad_out = torch.randn(N, Lq, M * D)
    value, loc, attn, grad_out = (t.to(dtype) for t in (value, loc, attn, grad_out))
    return shapes_t, lsi, value, loc, attn, grad_out


def run_reference(core, shapes_t, value, loc, attn, grad_out):
    v = value.clone().requires_grad_(True)
    lo = loc.clone().requires_grad_(True)
    a = attn.clone().requires_grad_(True)
    out = core(v, shapes_t, lo, a)
    gv, gl, ga = torch.autograd.grad(out, (v, lo, a), grad_out)
    return out.detach(), gv, gl, ga


CASES = {
    # name: (seed, shapes, N, M, D, Lq, P, dtype, loc_lo, loc_hi, value_scale, reftest)
    "reftest_f64": (3, [(6, 4), (3, 2)], 1, 2, 2, 2, 2, torch.float64, 0, 1, 1, True),
    "reftest_f32": (3, [(6, 4), (3, 2)], 1, 2, 2, 2, 2, torch.float32, 0, 1, 1, True),
    "d32_f32": (11, [(6, 20), (3, 10), (2, 5), (1, 3)], 2, 8, 32, 19, 4, torch.float32, -0.15, 1.15, 1, False),
    "d32_f64": (12, [(6, 20), (3, 10), (2, 5), (1, 3)], 1, 8, 32, 13, 4, torch.float64, -0.15, 1.15, 1, False),
    "d30_f64": (13, [(6, 4), (3, 2)], 1, 2, 30, 3, 2, torch.float64, -0.1, 1.1, 1, False),
    "d71_f64": (14, [(5, 7), (3, 2), (1, 1)], 2, 3, 71, 4, 3, torch.float64, -0.3, 1.3, 1, False),
    "d64_f32": (15, [(9, 11), (4, 5)], 2, 4, 64, 9, 4, torch.float32, -0.1, 1.1, 1, False),
    "d16_f32": (16, [(7, 9), (4, 5), (2, 3), (1, 2)], 2, 8, 16, 11, 4, torch.float32, -0.1, 1.1, 1, False),
}


def main():
    ref_shims.install()
    from lib.models.monodetr.ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch as core
    os.makedirs(OUT, exist_ok=True)
    for name, (seed, shapes, N, M, D, Lq, P, dtype, lo, hi, vs, rt) in CASES.items():
        shapes_t, lsi, value, loc, attn, grad_out = make_case(seed, shapes, N, M, D, Lq, P, dtype, lo, hi, vs, rt)
        out, gv, gl, ga = run_reference(core, shapes_t, value, loc, attn, grad_out)
        np.savez_compressed(
            os.path.join(OUT, f"msda_{name}.npz"),
            shapes=shapes_t.numpy(), lsi=lsi.numpy(), value=value.numpy(), loc=loc.numpy(), attn=attn.numpy(),
            grad_out=grad_out.numpy(), out=out.numpy(), grad_value=gv.numpy(), grad_loc=gl.numpy(),
            grad_attn=ga.numpy())
        print(name, tuple(value.shape), "->", tuple(out.shape))


if __name__ == "__main__":
    main()
