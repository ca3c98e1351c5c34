"""bench.py's extra legs (N = 1 only): SURVEY.md 8(f) rows timed on the device beside their CPU restatements, and the complete
training iteration in one CUDA graph.  Lives beside bench.py (not in the package) because the CPU legs execute oracle/ -- the
checker, used here only as the reported baseline; the device legs never touch it."""
import math
import os
import time

import numpy as np
import torch

from monodetr_b200 import _lib, build_monodetr
from monodetr_b200.ddp import FlatGradBucket
from monodetr_b200.monodetr import DEFAULT_MODEL_CFG
from monodetr_b200.bench_model import synthetic_batch

CRIT_CFG = {"num_classes": 3, "cls_loss_coef": 2, "focal_alpha": 0.25, "bbox_loss_coef": 5, "giou_loss_coef": 2, "3dcenter_loss_coef": 10,
            "dim_loss_coef": 1, "angle_loss_coef": 1, "depth_loss_coef": 1, "depth_map_loss_coef": 1, "set_cost_class": 2, "set_cost_bbox": 5,
            "set_cost_giou": 2, "set_cost_3dcenter": 10, "aux_loss": True, "dec_layers": 3}          # configs/monodetr.yaml:74-89


def synthetic_heads(seed, B, Q, C=3, n_aux=2, H=24, W=80):
    """Head outputs with MonoDETR.forward's shapes (monodetr.py:270-283), seeded."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)  # noqa: E731
    n = lambda *s: torch.randn(*s, generator=g)  # noqa: E731

    def heads():
        return {"pred_logits": n(B, Q, C) * 2 - 2, "pred_boxes": torch.cat([0.1 + 0.8 * r(B, Q, 2), 0.01 + 0.2 * r(B, Q, 4)], -1),
                "pred_3d_dim": 1.5 + 0.5 * n(B, Q, 3), "pred_depth": torch.stack([3 + 50 * r(B, Q), n(B, Q)], -1), "pred_angle": n(B, Q, 24)}
    out = heads()
    out["pred_depth_map_logits"] = n(B, 81, H, W)
    out["aux_outputs"] = [heads() for _ in range(n_aux)]
    return out


def synthetic_targets(seed, B, Gmax=50, max_gt=12, C=3):
    """The data loader's padded target arrays + validity mask (kitti_dataset.py:213-330), seeded."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)  # noqa: E731
    counts = torch.randint(1, max_gt + 1, (B,), generator=g)
    mask = torch.zeros(B, Gmax, dtype=torch.bool)
    for b in range(B):
        mask[b, torch.randperm(Gmax, generator=g)[:counts[b]]] = True
    ctr, lrtb = 0.05 + 0.9 * r(B, Gmax, 2), 0.01 + 0.25 * r(B, Gmax, 4)
    x0, y0, x1, y1 = ctr[..., 0] - lrtb[..., 0], ctr[..., 1] - lrtb[..., 2], ctr[..., 0] + lrtb[..., 1], ctr[..., 1] + lrtb[..., 3]
    return {"mask_2d": mask, "labels": torch.randint(0, C, (B, Gmax), generator=g).to(torch.int8),
            "boxes": torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, x1 - x0, y1 - y0], -1), "boxes_3d": torch.cat([ctr, lrtb], -1),
            "depth": 2 + 60 * r(B, Gmax, 1), "size_3d": 0.5 + 3 * r(B, Gmax, 3), "heading_bin": torch.randint(0, 12, (B, Gmax, 1), generator=g),
            "heading_res": (r(B, Gmax, 1) - 0.5) * (math.pi / 6)}


def synthetic_images(seed, sizes):
    g = np.random.default_rng(seed)
    return [g.integers(0, 256, (H, W, 3), dtype=np.uint8) for (W, H) in sizes]


def _timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def next_rows_probe(dev):
    """the steps either side of the hot path, each timed on the device (CUDA events, 10 calls) beside its CPU
    restatement on the host (oracle/, the checker -- used here only as the reported baseline).  B = 8, KITTI-shaped inputs."""
    from oracle import criterion as oc, decode as od, preprocess as op          # CPU baselines only
    from monodetr_b200.criterion import build_criterion
    from monodetr_b200.decode import decode_detections_device, extract_dets_from_outputs
    from monodetr_b200.preprocess import ImageBatchPreprocessor, get_affine_transform
    res = {}
    # f1: criterion (Hungarian matching + all losses), forward + backward of the weighted total
    out, padded = synthetic_heads(5, 8, 550), synthetic_targets(5, 8)
    crit = build_criterion(CRIT_CFG).to(dev).train()
    o = {k: (v.to(dev).requires_grad_(True) if torch.is_tensor(v) else [{kk: vv.to(dev).requires_grad_(True) for kk, vv in a.items()} for a in v])
         for k, v in out.items()}
    tg = {k: v.to(dev) for k, v in padded.items()}

    def crit_step():
        losses = crit(o, tg)
        sum(losses[k] * crit.weight_dict[k] for k in losses if k in crit.weight_dict).backward()
    def crit_step_fused():
        crit(o, tg)
        crit.weighted_sum().backward()
    n0 = _lib.launch_count()
    ms = _timed(crit_step)
    ms_fused = _timed(crit_step_fused)
    t0 = time.perf_counter()
    ro = {k: (v.clone().requires_grad_(True) if torch.is_tensor(v) else [{kk: vv.clone().requires_grad_(True) for kk, vv in a.items()} for a in v])
          for k, v in out.items()}
    rl, _ = oc.set_criterion(ro, padded, training=True)
    w = oc.weight_dict()
    sum(rl[k] * w[k] for k in rl if k in w).backward()
    res["criterion"] = {"ms_fwd_bwd": ms_fused, "ms_fwd_bwd_dict_sum": ms, "own_kernel_launches": (_lib.launch_count() - n0) // 26, "host_syncs": 0,
                        "cpu_port_ms": (time.perf_counter() - t0) * 1e3, "workload": "B=8, 550 queries x 3 decoder layers, 11 groups, <=12 objects/image; ms_fwd_bwd totals the loss with "
                                    "SetCriterion.weighted_sum(), ms_fwd_bwd_dict_sum with the reference trainer's Python sum over the dict"}
    # f3: post-process (top-k + decode), eval outputs of a batch of 32
    hh = synthetic_heads(3, 32, 50, n_aux=0)
    h = {"logits": hh["pred_logits"].numpy(), "boxes": hh["pred_boxes"].numpy(), "dim3": hh["pred_3d_dim"].numpy(),
         "depth": hh["pred_depth"].numpy(), "angle": hh["pred_angle"].numpy(),
         "img_size": np.tile(np.array([[1242.0, 375.0]], np.float32), (32, 1)),
         "P2": np.tile(np.array([[721.5377, 0, 609.5593, 44.85728], [0, 721.5377, 172.854, 0.2163791], [0, 0, 1, 0.002745884]], np.float32), (32, 1, 1)),
         "mean_size": np.array([[1.76255119, 0.66068622, 0.84422524], [1.52563191462, 1.62856739989, 3.88311640418],
                                [1.73698127, 0.59706367, 1.76282397]], np.float32)}
    ho = {k: v.to(dev) for k, v in hh.items() if torch.is_tensor(v)}
    sz, P2, ms3 = torch.from_numpy(h["img_size"]).to(dev), torch.from_numpy(h["P2"]).to(dev), torch.from_numpy(h["mean_size"]).to(dev)
    ms = _timed(lambda: decode_detections_device(extract_dets_from_outputs(ho, topk=50), sz, P2, ms3, 0.2))
    t0 = time.perf_counter()
    od.decode_dets(od.extract_dets(h["logits"], h["boxes"], h["dim3"], h["depth"], h["angle"], 50), h["img_size"], h["P2"], h["mean_size"], 0.2)
    res["postprocess"] = {"ms": ms, "cpu_port_ms": (time.perf_counter() - t0) * 1e3, "workload": "B=32, 50 queries x 3 classes, top-50, threshold 0.2"}
    # f4: warp + normalise 8 KITTI-sized frames to 1280x384 (sources already on the device)
    sizes = [(1242, 375), (1224, 370), (1238, 374), (1241, 376)] * 2
    imgs = synthetic_images(1, sizes)
    tinv = np.stack([get_affine_transform(np.array(s, np.float64) / 2, np.array(s, np.float64), 0, np.array([1280, 384]), inv=1)[1] for s in sizes])
    src = [torch.from_numpy(im).to(dev) for im in imgs]
    pre = ImageBatchPreprocessor((1280, 384), device=dev)
    ms = _timed(lambda: pre(src, tinv))
    nbytes = 8 * 3 * 384 * 1280 * 4 + sum(w * hh * 3 for w, hh in sizes)
    t0 = time.perf_counter()
    try:
        from PIL import Image
        for im, t in zip(imgs, tinv):
            x = np.array(Image.fromarray(im).transform((1280, 384), method=Image.AFFINE, data=tuple(t.reshape(-1).tolist()), resample=Image.BILINEAR))
            op.normalize(x)
        kind = "PIL + numpy (the reference's calls), 1 thread"
    except ImportError:
        for im, t in zip(imgs, tinv):
            op.preprocess(im, t.reshape(-1), (1280, 384))
        kind = "numpy port, 1 thread"
    res["preprocess"] = {"ms": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "cpu_ms": (time.perf_counter() - t0) * 1e3, "cpu_kind": kind,
                         "workload": "8 frames ~1242x375 u8 -> (8,3,384,1280) fp32"}
    return res


def full_train_step_probe(dev, B, steps, flush):
    """The COMPLETE training iteration of the reference's trainer (lib/helpers/trainer_helper.py:122-170) at the benchmark's shapes:
    model forward (train mode, dropout) -> SetCriterion (Hungarian matching + all losses, device-resident) -> weighted total ->
    backward -> fused AdamW, captured in ONE CUDA graph (nothing in it touches the host).  Reported next to the headline, which by
    SURVEY.md 8(d) uses the surrogate loss and no optimizer."""
    from monodetr_b200.criterion import build_criterion
    from monodetr_b200.optim import FusedAdamW
    torch.manual_seed(0)
    model, _ = build_monodetr(DEFAULT_MODEL_CFG)
    model = model.to(dev).train()
    crit = build_criterion(CRIT_CFG).to(dev).train()
    bucket = FlatGradBucket(model)
    opt = FusedAdamW(model, bucket, lr=2e-4, weight_decay=1e-4, device_step=True)
    images, calibs, sizes = (t.to(dev) for t in synthetic_batch(B, seed=77))
    tg = {k: v.to(dev) for k, v in synthetic_targets(77, B).items()}
    loss_buf = torch.zeros((), device=dev)

    def it():
        bucket.zero()
        out = model(images, calibs, None, sizes)
        crit(out, tg)
        total = crit.weighted_sum()
        total.backward()
        opt.step()
        loss_buf.copy_(total.detach())

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            it()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    l0 = float(loss_buf.item())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        it()
    bucket.freeze_sources()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        flush.zero_()
        a.record()
        graph.replay()
        b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    l1 = float(loss_buf.item())
    return {"value": B / (ms * 1e-3), "unit": "images/sec", "ms_per_step": ms, "loss_first": l0, "loss_last": l1,
            "config": f"batch {B}, 1280x384, train mode; forward + SetCriterion (device Hungarian matching, 26 loss terms) + backward + fused "
                      "AdamW (lr 2e-4) in one CUDA graph; synthetic padded targets (<= 12 objects / image); 256 MiB L2 flush between steps"}


