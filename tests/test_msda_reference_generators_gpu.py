"""The reference's own op test (lib/models/monodetr/ops/test.py:21-86) run against the drop-in boundary B1: the
top-level module `MultiScaleDeformableAttention` (the name ops/functions/ms_deform_attn_func.py:18 imports) serving
`ms_deform_attn_forward / ms_deform_attn_backward` from the sm_100a C ABI.

/root/reference is not present on the GPU box, so the generators of ops/test.py are RESTATED here with the same sizes,
seed, scalings and tolerances (N, M, D = 1, 2, 2; Lq, L, P = 2, 2, 2; shapes (6,4),(3,2); manual_seed(3); value * 0.01;
attention weights normalised over (level, point)); the autograd class below is the reference's glue
(ms_deform_attn_func.py:21-38) restated around the imported module, and `ms_deform_attn_core_pytorch` is served by its
pinned port oracle/msda_torch.py (tests/test_oracle_msda.py::test_torch_port_matches_reference).
"""
import pytest
import torch
from torch.autograd import Function, gradcheck
from torch.autograd.function import once_differentiable

from oracle.msda_torch import msda_core_torch

pytestmark = pytest.mark.gpu

N, M, D = 1, 2, 2
Lq, L, P = 2, 2, 2


class _RefGlue(Function):
    """ms_deform_attn_func.py:21-38: forward saves the five tensors, backward returns (gv, None, None, gl, ga, None)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step):
        import MultiScaleDeformableAttention as MSDA
        ctx.im2col_step = im2col_step
        output = MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                             attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        import MultiScaleDeformableAttention as MSDA
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        gv, gl, ga = MSDA.ms_deform_attn_backward(value, shapes, lsi, loc, attn, grad_output, ctx.im2col_step)
        return gv, None, None, gl, ga, None


def _setup():
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long).cuda()
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = sum((H * W).item() for H, W in shapes)
    return shapes, lsi, S


def _inputs(S, channels):
    value = torch.rand(N, S, M, channels).cuda() * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2).cuda()
    attn = torch.rand(N, Lq, M, L, P).cuda() + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value, loc, attn


def test_module_exports_the_pybind_names():
    import MultiScaleDeformableAttention as MSDA
    assert callable(MSDA.ms_deform_attn_forward) and callable(MSDA.ms_deform_attn_backward)      # vision.cpp:13-16
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):                          # ms_deform_attn.h:38
        shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
        lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
        MSDA.ms_deform_attn_forward(torch.rand(1, 30, 2, 2), shapes, lsi, torch.rand(1, 2, 2, 2, 2, 2), torch.rand(1, 2, 2, 2, 2), 2)


@torch.no_grad()
def test_check_forward_equal_with_pytorch_double():
    torch.manual_seed(3)
    shapes, lsi, S = _setup()
    value, loc, attn = _inputs(S, D)
    ref = msda_core_torch(value.double(), shapes, loc.double(), attn.double()).detach().cpu()
    out = _RefGlue.apply(value.double(), shapes, lsi, loc.double(), attn.double(), 2).detach().cpu()
    assert torch.allclose(out, ref)                                                                # test.py:40


@torch.no_grad()
def test_check_forward_equal_with_pytorch_float():
    torch.manual_seed(3)
    shapes, lsi, S = _setup()
    _inputs(S, D)                                       # (the reference draws the double-check inputs first)
    value, loc, attn = _inputs(S, D)
    ref = msda_core_torch(value, shapes, loc, attn).detach().cpu()
    out = _RefGlue.apply(value, shapes, lsi, loc, attn, 2).detach().cpu()
    assert torch.allclose(out, ref, rtol=1e-2, atol=1e-3)                                          # test.py:56
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-7)                                          # and the bar this repo holds


# test.py:84 lists [30, 32, 64, 71, 1025, 2048, 3096]; the full (slow-mode) Jacobians gradcheck materialises for the last two
# are 8 and 18 GB each, so they are left to the oracle comparison of the generic-D kernel (tests/test_msda_gpu.py, D = 1025).
@pytest.mark.parametrize("channels", [30, 32, 64, 71, 1025])
def test_check_gradient_numerical(channels):
    torch.manual_seed(3 + channels)
    shapes, lsi, S = _setup()
    value, loc, attn = _inputs(S, channels)
    value.requires_grad = True
    loc.requires_grad = True
    attn.requires_grad = True
    assert gradcheck(_RefGlue.apply, (value.double(), shapes, lsi, loc.double(), attn.double(), 2))   # test.py:78
