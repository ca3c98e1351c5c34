"""CPU: host logic of the drop-ins either side of the path (fused AdamW, post-process; the image pre-processor uploads to a real device
and stays with the -m gpu suite) and of
the op-extension boundary B1 (`MultiScaleDeformableAttention`), driven through the C-ABI-level stand-in for the device library
(tests/fake_device_lib.py) and compared with the oracle restatements pinned to the unmodified reference
(tests/test_oracle_{optim,decode,msda}.py).  What is under test is the HOST side: flat-buffer layout and grouping
rule of the optimizer, argument marshalling and result shaping of the reference-signature functions.  The kernels themselves
are pinned by the -m gpu suites.
"""
import numpy as np
import torch

import fake_device_lib          # tests/fake_device_lib.py (pytest puts this directory on sys.path)
from oracle import decode as od
from oracle import optim as oo


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(7, 5)
        self.b = torch.nn.Conv2d(3, 4, 3)
        self.norm = torch.nn.LayerNorm(5)
        self.label_enc = torch.nn.Embedding(4, 3)        # never receives a gradient: left untouched


def test_fused_adamw_flat_layout_and_update_rule(monkeypatch):
    from monodetr_b200 import optim
    fake_device_lib.install(monkeypatch)
    torch.manual_seed(0)
    model = _Toy()
    names = [n for n, p in model.named_parameters() if "label_enc" not in n]
    ref_p = {n: p.detach().clone() for n, p in model.named_parameters()}
    ref_m = {n: torch.zeros_like(p) for n, p in ref_p.items()}
    ref_v = {n: torch.zeros_like(p) for n, p in ref_p.items()}
    opt = optim.build_optimizer({"type": "adamw", "lr": 2e-3, "weight_decay": 1e-2}, model)
    assert [g["weight_decay"] for g in opt.param_groups] == [0, 1e-2]                   # optimizer_helper.py:9-16: 'bias' in name -> no decay
    assert all("bias" in n for n, p in model.named_parameters() if any(p is q for q in opt.param_groups[0]["params"]))
    for step in range(1, 4):
        grads = {n: torch.randn_like(p) for n, p in model.named_parameters() if n in names}
        for n, p in model.named_parameters():
            p.grad = grads[n].clone() if n in grads else None
        opt.step()
        oo.adamw_reference_step([ref_p[n] for n in names], [grads[n] for n in names], [ref_m[n] for n in names], [ref_v[n] for n in names],
                                step, 2e-3, 0.9, 0.999, 1e-8, [0.0 if "bias" in n else 1e-2 for n in names])
        for n, p in model.named_parameters():
            assert torch.allclose(p.detach(), ref_p[n], rtol=1e-6, atol=1e-7), (step, n)
    assert torch.equal(model.label_enc.weight.detach(), ref_p["label_enc.weight"])
    assert model.a.weight.data_ptr() >= opt.flat_p.data_ptr() and model.a.weight.data_ptr() < opt.flat_p.data_ptr() + opt.flat_p.numel() * 4


def test_post_process_reference_signature(monkeypatch):
    from monodetr_b200 import decode
    fake_device_lib.install(monkeypatch)
    h = od.synthetic_heads(3, 2, 50)
    outputs = {"pred_logits": torch.from_numpy(h["logits"]), "pred_boxes": torch.from_numpy(h["boxes"]), "pred_3d_dim": torch.from_numpy(h["dim3"]),
               "pred_depth": torch.from_numpy(h["depth"]), "pred_angle": torch.from_numpy(h["angle"])}
    dets = decode.extract_dets_from_outputs(outputs, K=50, topk=50)
    assert dets.shape == (2, 50, 37)
    want = od.decode_dets(dets.numpy(), h["img_size"], h["P2"], h["mean_size"], 0.2)

    class Calib:                                            # kitti_utils.py:136-155: the tester passes objects with .P2
        def __init__(self, P2):
            self.P2 = P2
    info = {"img_id": np.array([11, 42]), "img_size": h["img_size"]}
    res = decode.decode_detections(dets, info, [Calib(P) for P in h["P2"]], h["mean_size"], 0.2)
    assert list(res) == [11, 42]
    for i, img_id in enumerate([11, 42]):
        assert len(res[img_id]) == len(want[i])
        for got, ref in zip(res[img_id], want[i]):
            assert isinstance(got[0], int) and got[0] == ref[0]
            np.testing.assert_allclose(got[1:], ref[1:], rtol=1e-5, atol=1e-4)


def test_op_extension_boundary_names_and_contract(monkeypatch):
    """`MultiScaleDeformableAttention.ms_deform_attn_forward / _backward` (vision.cpp:13-16) hand the right pointers / sizes to the C
    ABI for fp32 and fp64, allocate their outputs, and return [grad_value, grad_loc, grad_attn]."""
    import MultiScaleDeformableAttention as MSDA
    from oracle import msda as oracle_msda
    fake_device_lib.install(monkeypatch)
    g = torch.Generator().manual_seed(1)
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    for dtype, tol in ((torch.float64, 1e-12), (torch.float32, 1e-5)):
        value = torch.randn(2, S, 3, 8, generator=g, dtype=dtype)
        loc = torch.rand(2, 5, 3, 2, 4, 2, generator=g, dtype=dtype) * 1.2 - 0.1
        attn = torch.softmax(torch.randn(2, 5, 3, 8, generator=g, dtype=dtype), -1).view(2, 5, 3, 2, 4)
        gout = torch.randn(2, 5, 24, generator=g, dtype=dtype)
        out = MSDA.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64)
        gv, gl, ga = MSDA.ms_deform_attn_backward(value, shapes, lsi, loc, attn, gout, 64)
        npv = [t.numpy() for t in (value, shapes, lsi, loc, attn)]
        np.testing.assert_allclose(out.numpy(), oracle_msda.msda_forward(*npv), rtol=tol, atol=tol)
        for got, want in zip((gv, gl, ga), oracle_msda.msda_backward(*npv, gout.numpy())):
            np.testing.assert_allclose(got.numpy(), want, rtol=10 * tol, atol=10 * tol)
