"""GPU parity of the fused elementwise chains (csrc/heads.cu) against the plain PyTorch expressions of the reference they
replace, forward and backward (autograd on the reference expression):
  box refinement  depthaware_transformer.py:602-613 + utils/misc.py:473-477
  query depth     monodetr.py:230-262
  depth tail      depth_predictor/depth_predictor.py:74-77, 93-104
  mean of 3 maps  depth_predictor.py:66;   sum_k mean(x_k^2): SURVEY.md 8(d) surrogate loss."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def _inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


@pytest.mark.parametrize("rd", [2, 6])
def test_box_refine(rd):
    from monodetr_b200 import functional as Fn
    g = torch.Generator(device="cuda").manual_seed(rd)
    tmp = torch.randn(3, 550, 6, device="cuda", generator=g, requires_grad=True)
    ref = torch.rand(3, 550, rd, device="cuda", generator=g)
    ref[0, :5] = torch.tensor([0.0, 1.0, 1e-6, 1 - 1e-6, 0.5, 0.3])[:rd].cuda()      # clamp corners
    ref.requires_grad_()
    dy = torch.randn(3, 550, 6, device="cuda", generator=g)
    y = Fn.box_refine(tmp, ref)
    gt, gr = torch.autograd.grad(y, (tmp, ref), dy)
    t2, r2 = tmp.detach().clone().requires_grad_(), ref.detach().clone().requires_grad_()
    if rd == 6:
        yr = (t2 + _inverse_sigmoid(r2)).sigmoid()
    else:
        yr = torch.cat((t2[..., :2] + _inverse_sigmoid(r2), t2[..., 2:]), -1).sigmoid()
    rt, rr = torch.autograd.grad(yr, (t2, r2), dy)
    assert _rel(y, yr) < 1e-6 and _rel(gt, rt) < 1e-5
    inner = (ref.detach() > 1e-4) & (ref.detach() < 1 - 1e-4)                          # away from the clamp kinks
    assert float(((gr - rr) * inner).abs().max()) < 1e-4 * float(rr.abs().max())
    # with a detached reference (every decoder layer but the first) no reference gradient is produced
    y2 = Fn.box_refine(tmp, ref.detach())
    assert torch.equal(y2, y)


@pytest.mark.parametrize("B,N,H,W", [(2, 50, 24, 80), (3, 550, 24, 80), (1, 7, 5, 9)])
def test_head_depth(B, N, H, W):
    from monodetr_b200 import functional as Fn
    g = torch.Generator(device="cuda").manual_seed(B * 10 + N)
    coord = torch.rand(B, N, 6, device="cuda", generator=g)
    coord[..., :2] = coord[..., :2] * 1.3 - 0.15                                      # some centres outside the map
    coord[0, 0, 4:] = 1e-4                                                            # clamp(min=1) active
    size3d = torch.randn(B, N, 3, device="cuda", generator=g)
    reg = torch.randn(B, N, 2, device="cuda", generator=g)
    wd = torch.rand(B, H, W, device="cuda", generator=g) * 60
    calibs = torch.zeros(B, 3, 4, device="cuda"); calibs[:, 0, 0] = 721.5377 + torch.arange(B, device="cuda")
    sizes = torch.tensor([[1242., 375.]], device="cuda").repeat(B, 1)
    dout = torch.randn(B, N, 2, device="cuda", generator=g)
    ins = [t.clone().requires_grad_() for t in (coord, size3d, reg, wd)]
    out = Fn.head_depth(*ins, calibs, sizes)
    grads = torch.autograd.grad(out, ins, dout)
    c, s3, rg, w = [t.clone().requires_grad_() for t in (coord, size3d, reg, wd)]
    hn = c[:, :, 4] + c[:, :, 5]
    h = torch.clamp(hn * sizes[:, 1:2], min=1.0)
    geo = s3[:, :, 0] / h * calibs[:, 0, 0].unsqueeze(1)
    centre = ((c[..., :2] - 0.5) * 2).detach()
    dm = F.grid_sample(w.unsqueeze(1), centre.unsqueeze(2), mode="bilinear", align_corners=True).squeeze(1)
    ref = torch.cat([((1. / (rg[:, :, 0:1].sigmoid() + 1e-6) - 1.) + geo.unsqueeze(-1) + dm) / 3, rg[:, :, 1:2]], -1)
    rgrads = torch.autograd.grad(ref, (c, s3, rg, w), dout)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)
    for a, b, name in zip(grads, rgrads, ("coord", "size3d", "reg", "wdepth")):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max())), name


@pytest.mark.parametrize("B,H,W", [(2, 24, 80), (1, 3, 5)])
def test_depth_tail(B, H, W):
    from monodetr_b200 import functional as Fn
    g = torch.Generator(device="cuda").manual_seed(B + H)
    nb, E, C, dmax = 81, 61, 256, 60.0
    logits = (torch.randn(B, H, W, nb, device="cuda", generator=g) * 3).requires_grad_()
    idx = torch.linspace(0, nb - 2, nb - 1, device="cuda")
    bin_size = 2 * (dmax - 1e-3) / ((nb - 1) * nb)
    bins = torch.cat(((idx + 0.5).pow(2) * bin_size / 2 - bin_size / 8 + 1e-3, torch.tensor([dmax], device="cuda")))
    emb = torch.randn(E, C, device="cuda", generator=g).requires_grad_()
    d_ip = torch.randn(B, H, W, C, device="cuda", generator=g)
    d_wd = torch.randn(B, H, W, device="cuda", generator=g)
    wd, ip = Fn.depth_tail(logits, bins, emb, dmax)
    gl, ge = torch.autograd.grad((wd, ip), (logits, emb), (d_wd, d_ip))
    l2, e2 = logits.detach().clone().requires_grad_(), emb.detach().clone().requires_grad_()
    p = F.softmax(l2, dim=-1)
    rwd = (p * bins.view(1, 1, 1, -1)).sum(-1)
    x = rwd.clamp(min=0, max=dmax)
    fl = x.floor()
    delta = (x - fl).unsqueeze(-1)
    fi = fl.long()
    ci = (fi + 1).clamp(max=E - 1)
    rip = F.embedding(fi, e2) * (1 - delta) + F.embedding(ci, e2) * delta
    rl, re = torch.autograd.grad((rwd, rip), (l2, e2), (d_wd, d_ip))
    assert _rel(wd, rwd) < 1e-5
    # a depth within float error of an integer may floor to the other side; the interpolated embedding is continuous there
    assert _rel(ip, rip) < 1e-4
    assert _rel(gl, rl) < 1e-4 and _rel(ge, re) < 1e-4
    # flat depth map (random-init model: every pixel in the same bin) -- the run-length path of the embedding gradient
    flat = torch.zeros(B, H, W, nb, device="cuda", requires_grad=True)
    wd2, ip2 = Fn.depth_tail(flat, bins, emb, dmax)
    (ge2,) = torch.autograd.grad(ip2, emb, d_ip)
    k = int(float(wd2.flatten()[0]))
    dl = float(wd2.flatten()[0]) - k
    expect = torch.zeros_like(emb)
    expect[k] = d_ip.sum((0, 1, 2)) * (1 - dl)
    expect[k + 1] = d_ip.sum((0, 1, 2)) * dl
    assert _rel(ge2, expect) < 1e-4


def test_mean3_and_sum_mean_squares():
    from monodetr_b200 import functional as Fn
    g = torch.Generator(device="cuda").manual_seed(0)
    a, b, c = (torch.randn(2, 24, 80, 256, device="cuda", generator=g, requires_grad=True) for _ in range(3))
    dy = torch.randn(2, 24, 80, 256, device="cuda", generator=g)
    y = Fn.mean3(a, b, c)
    ga, gb, gc = torch.autograd.grad(y, (a, b, c), dy)
    assert _rel(y, ((a + b + c) / 3).detach()) < 1e-6
    for gg in (ga, gb, gc):
        assert _rel(gg, dy / 3) < 1e-6
    xs = [torch.randn(s, device="cuda", generator=g, requires_grad=True) for s in [(8, 550, 3), (8, 550, 6), (8, 550, 24), (7,)]]
    xs.append(torch.randn(8, 24, 80, 81, device="cuda", generator=g).permute(0, 3, 1, 2).requires_grad_())   # non-contiguous view
    loss = Fn.sum_mean_squares(xs)
    grads = torch.autograd.grad(loss * 1.5, xs)
    ref = sum((x.detach() ** 2).mean() for x in xs)
    assert abs(float(loss) - float(ref)) < 1e-5 * float(ref)
    for x, gx in zip(xs, grads):
        assert gx.shape == x.shape and _rel(gx, 3.0 * x.detach() / x.numel()) < 1e-5
