"""GPU parity of the full model path (build_monodetr -> MonoDETR.forward / backward on the sm_100a kernels)
against (1) the golden fixtures generated from the UNMODIFIED reference and (2) the CPU oracle on the same seeded
inputs.  Tolerance: the north star's 1e-3, measured as max|a-b| / max|b| per output tensor.  Convolutions and
linears run on the TF32 tensor path (as the reference's own cuDNN path does under PyTorch defaults)."""
import os

import numpy as np
import pytest
import torch

from oracle import monodetr_torch as om

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
OUT_KEYS = ("pred_logits", "pred_boxes", "pred_3d_dim", "pred_depth", "pred_angle", "pred_depth_map_logits")
TOL = 1e-3


@pytest.fixture(autouse=True, params=["bf16x3", "tf32x3"])
def precision(request):
    """Both error-compensated tensor-core modes must hold the north star's 1e-3 on the whole model."""
    from monodetr_b200 import tc
    prev = tc.get_precision()
    tc.set_precision(request.param)
    yield request.param
    tc.set_precision(prev)


def _model(dropout=0.0):
    from monodetr_b200 import build_monodetr
    from monodetr_b200.monodetr import DEFAULT_MODEL_CFG
    cfg = dict(DEFAULT_MODEL_CFG, dropout=dropout)
    m, _ = build_monodetr(cfg)
    m.load_state_dict(om.with_aliases(om.deterministic_state_dict()))
    if dropout == 0.0:
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, torch.nn.MultiheadAttention):
                mod.dropout = 0.0
    return m.cuda()


def _rel(a, b):
    a = a.detach().float().cpu().numpy() if torch.is_tensor(a) else a
    b = b.detach().float().cpu().numpy() if torch.is_tensor(b) else b
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-20))


@pytest.mark.parametrize("name", ["model_eval_small", "model_eval_full", "model_train_full"])
def test_forward_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = _model(0.0)
    m.train(bool(g["training"]))
    images, calibs, sizes = om.synthetic_inputs(int(g["B"]), int(g["seed"]), H=int(g["H"]), W=int(g["W"]))
    with torch.no_grad():
        out = m(images.cuda(), calibs.cuda(), None, sizes.cuda())
    torch.cuda.synchronize()
    errs = {k: _rel(out[k], g[k]) for k in OUT_KEYS}
    for i, aux in enumerate(out["aux_outputs"]):
        for k, v in aux.items():
            errs[f"aux{i}_{k}"] = _rel(v, g[f"aux{i}_{k}"])
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs


def test_batch_independence_and_determinism():
    m = _model(0.0).eval()
    images, calibs, sizes = om.synthetic_inputs(3, 5)
    with torch.no_grad():
        o3 = m(images.cuda(), calibs.cuda(), None, sizes.cuda())
        o1 = m(images[1:2].cuda(), calibs[1:2].cuda(), None, sizes[1:2].cuda())
        o3b = m(images.cuda(), calibs.cuda(), None, sizes.cuda())
    for k in OUT_KEYS:
        assert torch.equal(o3[k], o3b[k]), k                              # forward is deterministic
        assert _rel(o3[k][1:2], o1[k]) < 1e-5, k                          # no cross-image coupling


def test_backward_matches_oracle_gradients():
    """train() shapes (550 queries, group self-attention), dropout 0, surrogate loss; gradients vs the CPU oracle."""
    m = _model(0.0).train()
    images, calibs, sizes = om.synthetic_inputs(1, 3, H=192, W=640)
    out = m(images.cuda(), calibs.cuda(), None, sizes.cuda())
    om.surrogate_loss(out).backward()
    torch.cuda.synchronize()
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in om.deterministic_state_dict().items()}
    ref_out = om.forward(sd, images, calibs, sizes, training=True)
    om.surrogate_loss(ref_out).backward()
    for k in OUT_KEYS:
        assert _rel(out[k], ref_out[k]) < TOL, k
    rels, worst = [], []
    for name, p in m.named_parameters():
        if name.startswith("depthaware_transformer.decoder.bbox_embed") or name.startswith("depthaware_transformer.decoder.dim_embed"):
            continue
        if not p.requires_grad:
            continue
        gref = sd[name].grad if name in sd else None
        if p.grad is None:
            assert gref is None or float(gref.abs().max()) == 0.0, f"{name}: missing gradient"
            continue
        assert gref is not None, name
        scale = float(gref.abs().max())
        if scale < 1e-6:
            continue
        r = float((p.grad.cpu() - gref).abs().max()) / scale
        rels.append(r)
        worst.append((r, name))
    worst.sort(reverse=True)
    print("worst gradient rel errs:", worst[:8], "median", sorted(rels)[len(rels) // 2])
    assert len(rels) > 250
    # Bars from measurement (tests/test_model_grad_gpu.py, profiles/r02_gradient_parity.txt): medians 3e-4..1e-3; the tail (1e-2, query_embed 4.5e-2 through the decoder's sampling offsets) is ReLU / max-pool / sampling-cell selections flipping on forward noise, which an fp32
    # CPU run shows against fp64 as well (5.5e-3).
    assert sorted(rels)[len(rels) // 2] < 3e-3
    assert worst[0][0] < 0.1, worst[:5]


def test_train_step_with_dropout_runs_and_is_finite():
    m = _model(0.1).train()
    images, calibs, sizes = om.synthetic_inputs(2, 1)
    out = m(images.cuda(), calibs.cuda(), None, sizes.cuda())
    loss = om.surrogate_loss(out)
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    n = 0
    for name, p in m.named_parameters():
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), name
            n += 1
    assert n == 313                                                        # gradient-receiving tensors


@pytest.mark.parametrize("B,H,W,N", [(2, 24, 80, 50), (3, 6, 20, 550), (1, 5, 7, 33)])
def test_depth_sample_matches_grid_sample(B, H, W, N):
    """mdb_depth_sample_* against F.grid_sample(bilinear, zeros, align_corners=True) as called at monodetr.py:248-253,
    with centres beyond [-1, 1] to exercise the zero padding; gradient wrt the map (the centres are detached)."""
    import torch.nn.functional as F
    from monodetr_b200 import functional as Fn
    g = torch.Generator(device="cuda").manual_seed(B * 100 + N)
    depth = torch.rand(B, H, W, device="cuda", generator=g).requires_grad_()
    xy = torch.rand(B, N, 2, device="cuda", generator=g) * 2.4 - 1.2
    xy[0, 0] = torch.tensor([-1.0, -1.0]); xy[0, 1] = torch.tensor([1.0, 1.0])
    dout = torch.randn(B, N, device="cuda", generator=g)
    out = Fn.depth_sample(depth, xy)
    (gd,) = torch.autograd.grad(out, depth, dout)
    d2 = depth.detach().clone().requires_grad_()
    ref = F.grid_sample(d2[:, None], xy[:, :, None], mode="bilinear", align_corners=True).squeeze(1).squeeze(-1)
    (rd,) = torch.autograd.grad(ref, d2, dout)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6)
    assert torch.allclose(gd, rd, rtol=1e-5, atol=1e-5)
