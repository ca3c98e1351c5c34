"""GPU parity of the fused AdamW (csrc/optim.cu, monodetr_b200/optim.py) against the reference's update
(lib/helpers/optimizer_helper.py:88-127) restated in oracle/optim.py with the same torch operations, run on the SAME device.
The kernel reproduces torch's per-operation rounding (explicit FMAs where torch's kernels contract, separate roundings
elsewhere), so the comparison is held to a few ulp: rtol 4e-7 on parameters and both moment buffers."""
import pytest
import torch

from oracle.optim import adamw_reference_step

pytestmark = pytest.mark.gpu


def _toy():
    torch.manual_seed(0)
    m = torch.nn.ModuleDict({
        "a": torch.nn.Linear(37, 19),            # odd sizes: alignment padding between the tensors of the flat buffer
        "norm": torch.nn.LayerNorm(19),
        "b": torch.nn.Linear(19, 3),
        "emb": torch.nn.Embedding(11, 5),
    })
    return m.cuda()


def _check(params, names, opt, steps, lr, wd):
    ref = [p.detach().clone() for p in params]
    ms = [torch.zeros_like(p) for p in params]
    vs = [torch.zeros_like(p) for p in params]
    wds = [0.0 if "bias" in n else wd for n in names]
    g = torch.Generator(device="cuda").manual_seed(1)
    exact = True
    for step in range(1, steps + 1):
        grads = [torch.randn(p.shape, device="cuda", generator=g) * (0.5 + step) for p in params]
        for p, gr in zip(params, grads):
            p.grad = gr.clone()
        opt.step()
        adamw_reference_step(ref, grads, ms, vs, step, lr, 0.9, 0.999, 1e-8, wds)
        for p, r, n in zip(params, ref, names):
            assert torch.allclose(p.detach(), r, rtol=4e-7, atol=1e-9), (n, step, float((p.detach() - r).abs().max()))
            exact = exact and torch.equal(p.detach(), r)
    for off, m_ref, v_ref in zip(opt.bucket.offsets, ms, vs):
        n = m_ref.numel()
        assert torch.allclose(opt.exp_avg[off:off + n].view_as(m_ref), m_ref, rtol=4e-7, atol=1e-12)
        assert torch.allclose(opt.exp_avg_sq[off:off + n].view_as(v_ref), v_ref, rtol=4e-7, atol=1e-12)
    used = torch.zeros(opt.bucket.numel, dtype=torch.bool, device="cuda")
    for off, p in zip(opt.bucket.offsets, params):
        used[off:off + p.numel()] = True
    assert not opt.flat_p[~used].any() and not opt.exp_avg[~used].any() and not opt.exp_avg_sq[~used].any()     # padding stays zero
    print("bit-exact vs torch's kernels:", exact)


@pytest.mark.parametrize("device_step", [False, True])
def test_fused_adamw_matches_reference_update(device_step):
    from monodetr_b200.ddp import FlatGradBucket
    from monodetr_b200.optim import FusedAdamW
    model = _toy()
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    bucket = FlatGradBucket(model)
    opt = FusedAdamW(model, bucket, lr=2e-4, weight_decay=1e-4, device_step=device_step)
    for n, p in model.named_parameters():                    # flattening moved the storage, not the values
        assert torch.equal(p.detach(), before[n])
        assert opt.flat_p.data_ptr() <= p.data_ptr() < opt.flat_p.data_ptr() + opt.flat_p.numel() * 4
    assert bucket.names[:bucket.names.index("a.bias")] == [n for n in bucket.names if "bias" not in n]   # decay tensors first
    assert bucket.n_decay == bucket.offsets[bucket.names.index("a.bias")] and all(o % 32 == 0 for o in bucket.offsets)
    assert all(p.data_ptr() % 128 == 0 for p in bucket.params)
    _check(bucket.params, bucket.names, opt, 6, 2e-4, 1e-4)


def test_fused_adamw_on_the_model():
    from monodetr_b200 import build_monodetr
    from monodetr_b200.monodetr import DEFAULT_MODEL_CFG
    from monodetr_b200.optim import build_optimizer
    torch.manual_seed(0)
    model, _ = build_monodetr(DEFAULT_MODEL_CFG)
    model = model.cuda()
    keys = list(model.state_dict().keys())
    opt = build_optimizer({"type": "adamw", "lr": 2e-4, "weight_decay": 1e-4}, model)
    assert list(model.state_dict().keys()) == keys
    b = opt.bucket
    assert len(b.params) == 313 and b.param_numel == 37056453  # gradient-receiving tensors
    untouched = {n: p.detach().clone() for n, p in model.named_parameters() if all(p is not q for q in b.params)}
    _check(b.params, b.names, opt, 2, 2e-4, 1e-4)
    for n, p in model.named_parameters():                       # never-used / frozen tensors are not updated (optimizer_helper.py:95-96)
        if n in untouched:
            assert torch.equal(p.detach(), untouched[n]), n
    # the model still runs on the flattened parameters (every tensor 128-byte aligned) and trains
    from oracle import monodetr_torch as om
    images, calibs, sizes = om.synthetic_inputs(1, 3, H=192, W=640)
    model.train()
    for _ in range(2):
        opt.zero_grad()
        out = model(images.cuda(), calibs.cuda(), None, sizes.cuda())
        loss = om.surrogate_loss(out)
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    assert torch.isfinite(loss) and all(torch.isfinite(p).all() for p in b.params)
