"""CPU: the HOST logic of the one-node encoder layer (monodetr_b200.functional._EncoderLayer) -- which kernel is fed what, which
gradient goes where, which fan-in sums ride in which GEMM epilogue -- checked with every kernel wrapper replaced by a plain
torch restatement of what that kernel computes (so nothing here needs a GPU or the library's compute paths), against the
oracle's statement of the layer (oracle/monodetr_torch.py: depthaware_transformer.py:315-354 of the reference) under autograd.
The kernels themselves are pinned by the -m gpu suites; this test pins the chain rule written by hand around them.
"""
import torch
import torch.nn.functional as F

from oracle import monodetr_torch as om
from oracle.msda_torch import msda_core_torch


def _prep(off, logits, ref, shapes, M):
    B, Lq = off.shape[:2]
    off = off.view(B, Lq, M, 4, 4, 2)
    attn = F.softmax(logits.view(B, Lq, M, 16), -1).view(B, Lq, M, 4, 4)
    norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1).to(off.dtype)
    loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    return loc, attn


def _install_torch_kernels(monkeypatch):
    from monodetr_b200 import functional as Fn, kernels as K, tc

    monkeypatch.setattr(tc, "lookup_split", lambda w: w.detach())
    monkeypatch.setattr(tc, "get_precision", lambda: "bf16x3")

    def linear_forward(x, w, bias=None, residual=None, relu=False, round_out=False):
        y = x @ w.t()
        if bias is not None:
            y = y + bias
        if residual is not None:
            y = y + residual
        return torch.relu(y) if relu else y

    def linear_dgrad(dy, w, residual=None, relu_mask=None):
        dx = dy @ w
        if residual is not None:
            dx = dx + residual
        if relu_mask is not None:
            dx = dx * (relu_mask > 0)
        return dx

    def linear_wgrad(dy, x, with_bias_grad=False):
        return (dy.t() @ x, dy.sum(0)) if with_bias_grad else dy.t() @ x

    def ln_fwd(x, res, gamma, beta, eps=1e-5, drop_p=0.0, site=0, seed=None):
        assert drop_p == 0.0
        z = x + res
        mean = z.mean(-1)
        rstd = (z.var(-1, unbiased=False) + eps).rsqrt()
        return (z - mean[:, None]) * rstd[:, None] * gamma + beta, mean, rstd

    def ln_bwd(dy, x, res, gamma, mean, rstd, drop_p=0.0, site=0, seed=None):
        with torch.enable_grad():                       # (a once_differentiable backward runs under no_grad)
            z = (x + res).detach().requires_grad_()
            g = gamma.detach().requires_grad_()
            b = torch.zeros_like(gamma).requires_grad_()
            y = F.layer_norm(z, (z.shape[-1],), g, b)
            dz, dg, db = torch.autograd.grad(y, (z, g, b), dy)
        return dz, dz, dg, db

    def msda_fwd(value, shapes, lsi, off, logits, refc):
        loc, attn = _prep(off, logits, refc, shapes, value.shape[2])
        return msda_core_torch(value, shapes, loc, attn)

    def msda_bwd(value, shapes, lsi, off, logits, refc, dout):
        with torch.enable_grad():
            v, o, l = (t.detach().requires_grad_() for t in (value, off, logits))
            loc, attn = _prep(o, l, refc, shapes, value.shape[2])
            return torch.autograd.grad(msda_core_torch(v, shapes, loc, attn), (v, o, l), dout)

    monkeypatch.setattr(tc, "linear_forward", linear_forward)
    monkeypatch.setattr(tc, "linear_dgrad", linear_dgrad)
    monkeypatch.setattr(tc, "linear_wgrad", linear_wgrad)
    monkeypatch.setattr(K, "add_layernorm_forward", ln_fwd)
    monkeypatch.setattr(K, "add_layernorm_backward", ln_bwd)
    monkeypatch.setattr(Fn, "msda_fused_forward_raw", msda_fwd)
    monkeypatch.setattr(Fn, "msda_fused_backward_raw", msda_bwd)
    return Fn


def test_one_node_encoder_layer_forward_and_every_gradient(monkeypatch):
    Fn = _install_torch_kernels(monkeypatch)
    from monodetr_b200.depthaware_transformer import VisualEncoderLayer
    torch.manual_seed(0)
    dt = torch.float64
    layer = VisualEncoderLayer(256, 256, 0.1, "relu", 4, 8, 4).to(dt).eval()       # eval: dropout off, gradients still flow
    for prm in layer.parameters():                                               # away from the zero-initialised offsets weight
        prm.data.add_(torch.randn_like(prm) * 0.05)
    shapes = [(6, 10), (3, 5), (2, 3), (1, 2)]
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    S = int(shapes_t.prod(1).sum())
    B = 2
    src = torch.randn(B, S, 256, dtype=dt, requires_grad=True)
    pos = torch.randn(S, 256, dtype=dt, requires_grad=True)
    ref = om.encoder_reference_points(shapes, B, "cpu").to(dt)
    dout = torch.randn(B, S, 256, dtype=dt)

    out = Fn.encoder_layer(layer, src, pos, ref, shapes_t, lsi)
    params = list(layer.parameters())
    grads = torch.autograd.grad(out, [src, pos] + params, dout)

    # the oracle's statement of the same layer, differentiated by autograd
    sd = {"e." + k: v for k, v in layer.state_dict(keep_vars=True).items()}
    src2 = om.ms_deform_attn(sd, "e.self_attn", src + pos, ref, src, shapes_t)
    mem = om.layer_norm(sd, "e.norm1", src + src2)
    want = om.layer_norm(sd, "e.norm2", mem + om.linear(sd, "e.linear2", F.relu(om.linear(sd, "e.linear1", mem))))
    want_grads = torch.autograd.grad(want, [src, pos] + params, dout)

    assert torch.allclose(out, want, rtol=1e-10, atol=1e-10)
    names = ["src", "pos"] + [n for n, _ in layer.named_parameters()]
    for n, g, w in zip(names, grads, want_grads):
        assert g.shape == w.shape, n
        assert torch.allclose(g, w, rtol=1e-8, atol=1e-9), (n, float((g - w).abs().max()))


def test_fusable_only_on_the_configuration_the_node_covers():
    from monodetr_b200 import functional as Fn
    from monodetr_b200.depthaware_transformer import VisualEncoderLayer
    layer = VisualEncoderLayer(256, 256, 0.1, "relu", 4, 8, 4)
    src = torch.zeros(1, 10, 256)
    ref = torch.zeros(1, 10, 4, 2)
    assert not Fn.encoder_layer_fusable(layer, src, ref, None)                    # CPU tensor: the separate nodes (which raise loudly)
