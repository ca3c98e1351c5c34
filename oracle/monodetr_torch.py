"""oracle/monodetr_torch.py -- TEST / BASELINE INFRASTRUCTURE (never imported by monodetr_b200/).

Plain-PyTorch fp32 restatement of the reference's MonoDETR forward pass as PURE FUNCTIONS over a
state_dict (same keys/shapes as the reference model), runnable on CPU with all host threads.
It is (1) the checker for the sm_100a path in tests/ and smoke(), and (2) the CPU baseline / `--impl
reference` arm of bench.py (the Python reference itself cannot travel to the GPU box).

Each function cites the reference code it restates (paths relative to /root/reference):
  backbone            lib/models/monodetr/backbone.py:27-127 + torchvision resnet50 v1.5 (un-vendored; restated)
  position embedding  lib/models/monodetr/position_encoding.py:20-56
  input_proj          lib/models/monodetr/monodetr.py:83-91,156-178
  depth predictor     lib/models/monodetr/depth_predictor/depth_predictor.py:56-104, transformer.py:16-65
  transformer         lib/models/monodetr/depthaware_transformer.py:199-312 (+ :315-384 encoder, :387-626 decoder)
  MSDeformAttn        lib/models/monodetr/ops/modules/ms_deform_attn.py:122-162, core = oracle/msda_torch.py
  heads               lib/models/monodetr/monodetr.py:212-283
Pinned by tests/test_oracle_model.py against the UNMODIFIED reference imported from /root/reference
(same state_dict, same inputs) and by the committed fixture tests/golden/model_*.npz.
Dropout is the identity here (parity runs use eval() / dropout 0; the timing baseline keeps train-mode shapes).
"""
import math
import os
import time

import torch
import torch.nn.functional as F

from .msda_torch import msda_core_torch

CFG = dict(num_classes=3, hidden_dim=256, nheads=8, enc_layers=3, dec_layers=3, dim_feedforward=256, num_queries=50,
           group_num=11, num_feature_levels=4, enc_n_points=4, dec_n_points=4, num_depth_bins=80, depth_min=1e-3,
           depth_max=60.0)

RESNET_LAYERS = [("layer1", 64, 3, 1), ("layer2", 128, 4, 2), ("layer3", 256, 6, 2), ("layer4", 512, 3, 2)]

# Test switches (both default off = the reference's arithmetic, which is what the golden / reference pins check):
#   FREEZE_SAMPLING  treat the MSDeformAttn sampling locations as constants in backward.  d(bilinear)/d(location) is
#                    discontinuous at cell borders, so gradients THROUGH the locations amplify 1e-6 forward noise arbitrarily;
#                    with the locations frozen every remaining gradient is a smooth function and can be compared tightly.
#   DROPOUT_P        apply dropout with torch's RNG at the reference's 34 sites (depthaware_transformer.py:341-349,456-513,
#                    depth_predictor/transformer.py:59-65) for STATISTICAL comparisons with the sm_100a path's hash masks.
FREEZE_SAMPLING = False
DROPOUT_P = 0.0


def _drop(x):
    return F.dropout(x, DROPOUT_P, training=True) if DROPOUT_P > 0 else x


def inverse_sigmoid(x, eps=1e-5):       # utils/misc.py:473-477
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def frozen_bn(sd, p, x):                # backbone.py:54-64
    scale = sd[p + ".weight"] * (sd[p + ".running_var"] + 1e-5).rsqrt()
    bias = sd[p + ".bias"] - sd[p + ".running_mean"] * scale
    return x * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


def bottleneck(sd, p, x, stride):       # torchvision Bottleneck v1.5 (stride on the 3x3)
    out = F.relu(frozen_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    out = F.relu(frozen_bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1)))
    out = frozen_bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]))
    if (p + ".downsample.0.weight") in sd:
        x = frozen_bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    return F.relu(out + x)


def backbone(sd, images):               # backbone.py:75-90 (returns layer2, layer3, layer4)
    p = "backbone.0.body."
    x = F.relu(frozen_bn(sd, p + "bn1", F.conv2d(images, sd[p + "conv1.weight"], stride=2, padding=3)))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for name, _, blocks, stride in RESNET_LAYERS:
        for b in range(blocks):
            x = bottleneck(sd, f"{p}{name}.{b}", x, stride if b == 0 else 1)
        if name != "layer1":
            feats.append(x)
    return feats


def position_embedding_sine(B, H, W, device, num_pos_feats=128, temperature=10000):   # position_encoding.py:36-56, mask all-False
    scale = 2 * math.pi
    y_embed = torch.arange(1, H + 1, dtype=torch.float32, device=device).view(1, H, 1).expand(B, H, W)
    x_embed = torch.arange(1, W + 1, dtype=torch.float32, device=device).view(1, 1, W).expand(B, H, W)
    y_embed = y_embed / (H + 1e-6) * scale
    x_embed = x_embed / (W + 1e-6) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=device)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def conv_gn(sd, p, x, stride=1, padding=0, relu=False):     # Sequential(Conv2d, GroupNorm(32, C))
    x = F.conv2d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], stride=stride, padding=padding)
    x = F.group_norm(x, 32, sd[p + ".1.weight"], sd[p + ".1.bias"])
    return F.relu(x) if relu else x


def mha(sd, p, q, k, v, nheads=8):      # nn.MultiheadAttention forward, seq-first (L, B, E); returns attn output only
    out, _ = F.multi_head_attention_forward(
        q, k, v, q.shape[-1], nheads, sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"], None, None, False, DROPOUT_P,
        sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"], training=DROPOUT_P > 0, need_weights=False)
    return out


def layer_norm(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"])


def linear(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def mlp(sd, p, x, num_layers):          # monodetr.py:535-547
    for i in range(num_layers):
        x = linear(sd, f"{p}.layers.{i}", x)
        if i < num_layers - 1:
            x = F.relu(x)
    return x


def ms_deform_attn(sd, p, query, reference_points, input_flatten, spatial_shapes, n_heads=8, n_levels=4, n_points=4):
    """ops/modules/ms_deform_attn.py:122-162 (no padding mask on this path: masks are all-False)."""
    N, Lq, _ = query.shape
    S = input_flatten.shape[1]
    value = linear(sd, p + ".value_proj", input_flatten).view(N, S, n_heads, -1)
    off = linear(sd, p + ".sampling_offsets", query).view(N, Lq, n_heads, n_levels, n_points, 2)
    attn = linear(sd, p + ".attention_weights", query).view(N, Lq, n_heads, n_levels * n_points)
    attn = F.softmax(attn, -1).view(N, Lq, n_heads, n_levels, n_points)
    if reference_points.shape[-1] == 2:
        normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
        loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    else:   # 6-d (cx, cy, l, r, t, b): :154-155
        loc = reference_points[:, :, None, :, None, :2] + off / n_points * (
            reference_points[:, :, None, :, None, 2::2] + reference_points[:, :, None, :, None, 3::2]) * 0.5
    if FREEZE_SAMPLING:
        loc = loc.detach()
    out = msda_core_torch(value, spatial_shapes, loc, attn)
    return linear(sd, p + ".output_proj", out)


def depth_predictor(sd, srcs, pos1, cfg=CFG):           # depth_predictor.py:56-104
    p = "depth_predictor."
    src_16 = conv_gn(sd, p + "proj", srcs[1])
    src_32 = conv_gn(sd, p + "upsample", F.interpolate(srcs[2], size=src_16.shape[-2:], mode="bilinear"))
    src_8 = conv_gn(sd, p + "downsample", srcs[0], stride=2, padding=1)
    src = (src_8 + src_16 + src_32) / 3
    h = p + "depth_head"
    src = F.relu(F.group_norm(F.conv2d(src, sd[h + ".0.weight"], sd[h + ".0.bias"], padding=1), 32, sd[h + ".1.weight"], sd[h + ".1.bias"]))
    src = F.relu(F.group_norm(F.conv2d(src, sd[h + ".3.weight"], sd[h + ".3.bias"], padding=1), 32, sd[h + ".4.weight"], sd[h + ".4.bias"]))
    depth_logits = F.conv2d(src, sd[p + "depth_classifier.weight"], sd[p + "depth_classifier.bias"])
    depth_probs = F.softmax(depth_logits, dim=1)
    weighted_depth = (depth_probs * sd[p + "depth_bin_values"].reshape(1, -1, 1, 1)).sum(dim=1)
    B, C, H, W = src.shape
    s = src.flatten(2).permute(2, 0, 1)
    pos = pos1.flatten(2).permute(2, 0, 1)
    e = p + "depth_encoder.layers.0"
    qk = s + pos
    s = layer_norm(sd, e + ".norm1", s + _drop(mha(sd, e + ".self_attn", qk, qk, s)))
    s = layer_norm(sd, e + ".norm2", s + _drop(linear(sd, e + ".linear2", _drop(F.relu(linear(sd, e + ".linear1", s))))))
    depth_embed = s.permute(1, 2, 0).reshape(B, C, H, W)
    d = weighted_depth.clamp(min=0, max=cfg["depth_max"])
    table = sd[p + "depth_pos_embed.weight"]
    fl = d.floor()
    delta = (d - fl).unsqueeze(-1)
    fl = fl.long()
    ce = (fl + 1).clamp(max=table.shape[0] - 1)
    ip = (table[fl] * (1 - delta) + table[ce] * delta).permute(0, 3, 1, 2)
    return depth_logits, depth_embed + ip, weighted_depth, ip


def encoder_reference_points(shapes, B, device):        # depthaware_transformer.py:363-376 with valid_ratios == 1
    pts = []
    for (H, W) in shapes:
        ry, rx = torch.meshgrid(torch.linspace(0.5, H - 0.5, H, dtype=torch.float32, device=device),
                                torch.linspace(0.5, W - 0.5, W, dtype=torch.float32, device=device), indexing="ij")
        pts.append(torch.stack((rx.reshape(-1) / W, ry.reshape(-1) / H), -1))
    ref = torch.cat(pts, 0)[None].expand(B, -1, -1)
    return ref[:, :, None].expand(-1, -1, len(shapes), -1)


def transformer(sd, srcs, pos_embeds, query_embed, depth_pos_embed, training, cfg=CFG):
    """depthaware_transformer.py:199-312 default branch (not two_stage / dab / dino)."""
    p = "depthaware_transformer."
    B = srcs[0].shape[0]
    dev = srcs[0].device
    shapes = [tuple(s.shape[-2:]) for s in srcs]
    src_flatten = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    lvl_pos = torch.cat([pe.flatten(2).transpose(1, 2) + sd[p + "level_embed"][l].view(1, 1, -1)
                         for l, pe in enumerate(pos_embeds)], 1)
    spatial_shapes = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    ref_enc = encoder_reference_points(shapes, B, dev)
    memory = src_flatten
    for l in range(cfg["enc_layers"]):          # :345-354
        e = f"{p}encoder.layers.{l}"
        src2 = ms_deform_attn(sd, e + ".self_attn", memory + lvl_pos, ref_enc, memory, spatial_shapes)
        memory = layer_norm(sd, e + ".norm1", memory + _drop(src2))
        memory = layer_norm(sd, e + ".norm2", memory + _drop(linear(sd, e + ".linear2", _drop(F.relu(linear(sd, e + ".linear1", memory))))))
    c = memory.shape[-1]
    query_pos, tgt = torch.split(query_embed, c, dim=1)     # :283-287
    query_pos = query_pos.unsqueeze(0).expand(B, -1, -1)
    tgt = tgt.unsqueeze(0).expand(B, -1, -1)
    reference_points = linear(sd, p + "reference_points", query_pos).sigmoid()
    init_reference = reference_points
    dpe = depth_pos_embed.flatten(2).permute(2, 0, 1)       # (1920, B, C)
    output = tgt
    G = cfg["group_num"]
    inter, inter_ref, inter_dim = [], [], []
    for l in range(cfg["dec_layers"]):                      # :548-626 / :437-515
        d = f"{p}decoder.layers.{l}"
        ref_in = reference_points[:, :, None].expand(-1, -1, len(shapes), -1)     # valid_ratios == 1
        tgt2 = mha(sd, d + ".cross_attn_depth", output.transpose(0, 1), dpe, dpe).transpose(0, 1)
        t = layer_norm(sd, d + ".norm_depth", output + _drop(tgt2))
        qk = t + query_pos
        q = (linear(sd, d + ".sa_qcontent_proj", qk) + linear(sd, d + ".sa_qpos_proj", qk)).transpose(0, 1)
        k = (linear(sd, d + ".sa_kcontent_proj", qk) + linear(sd, d + ".sa_kpos_proj", qk)).transpose(0, 1)
        v = t.transpose(0, 1)                               # :477 (sa_v_proj output is discarded)
        nq = q.shape[0]
        if training:                                        # :480-494 group fold (num_noise = 0)
            q = torch.cat(q.split(nq // G, dim=0), dim=1)
            k = torch.cat(k.split(nq // G, dim=0), dim=1)
            v = torch.cat(v.split(nq // G, dim=0), dim=1)
        tgt2 = mha(sd, d + ".self_attn", q, k, v)
        tgt2 = torch.cat(tgt2.split(B, dim=1), dim=0).transpose(0, 1) if training else tgt2.transpose(0, 1)
        t = layer_norm(sd, d + ".norm2", t + _drop(tgt2))
        tgt2 = ms_deform_attn(sd, d + ".cross_attn", t + query_pos, ref_in, memory, spatial_shapes)
        t = layer_norm(sd, d + ".norm1", t + _drop(tgt2))
        output = layer_norm(sd, d + ".norm3", t + _drop(linear(sd, d + ".linear2", _drop(F.relu(linear(sd, d + ".linear1", t))))))
        tmp = mlp(sd, f"bbox_embed.{l}", output, 3)         # :602-613
        if reference_points.shape[-1] == 6:
            new_ref = (tmp + inverse_sigmoid(reference_points)).sigmoid()
        else:
            new_ref = torch.cat((tmp[..., :2] + inverse_sigmoid(reference_points), tmp[..., 2:]), -1).sigmoid()
        reference_points = new_ref.detach()
        inter.append(output)
        inter_ref.append(reference_points)
        inter_dim.append(mlp(sd, f"dim_embed_3d.{l}", output, 2))
    return torch.stack(inter), init_reference, torch.stack(inter_ref), torch.stack(inter_dim)


def forward(sd, images, calibs, img_sizes, training=False, cfg=CFG):
    """monodetr.py:150-283.  Returns the reference's output dict."""
    feats = backbone(sd, images)
    B = images.shape[0]
    srcs = [conv_gn(sd, f"input_proj.{l}", f) for l, f in enumerate(feats)]
    srcs.append(conv_gn(sd, "input_proj.3", feats[-1], stride=2, padding=1))
    pos = [position_embedding_sine(B, s.shape[2], s.shape[3], s.device) for s in srcs]
    nq = cfg["num_queries"] * (cfg["group_num"] if training else 1)
    query_embeds = sd["query_embed.weight"][:nq]
    depth_logits, depth_pos_embed, weighted_depth, _ = depth_predictor(sd, srcs, pos[1], cfg)
    hs, init_ref, inter_refs, inter_dims = transformer(sd, srcs, pos, query_embeds, depth_pos_embed, training, cfg)
    coords, classes, dims, depths, angles = [], [], [], [], []
    for lvl in range(hs.shape[0]):
        reference = inverse_sigmoid(init_ref if lvl == 0 else inter_refs[lvl - 1])
        tmp = mlp(sd, f"bbox_embed.{lvl}", hs[lvl], 3)
        if reference.shape[-1] == 6:
            tmp = tmp + reference
        else:
            tmp = torch.cat((tmp[..., :2] + reference, tmp[..., 2:]), -1)
        coord = tmp.sigmoid()
        coords.append(coord)
        classes.append(linear(sd, f"class_embed.{lvl}", hs[lvl]))
        size3d = inter_dims[lvl]
        dims.append(size3d)
        box_h = torch.clamp((coord[:, :, 4] + coord[:, :, 5]) * img_sizes[:, 1:2], min=1.0)
        depth_geo = size3d[:, :, 0] / box_h * calibs[:, 0, 0].unsqueeze(1)
        depth_reg = mlp(sd, f"depth_embed.{lvl}", hs[lvl], 2)
        centre = ((coord[..., :2] - 0.5) * 2).unsqueeze(2).detach()
        depth_map = F.grid_sample(weighted_depth.unsqueeze(1), centre, mode="bilinear", align_corners=True).squeeze(1)
        depths.append(torch.cat([((1. / (depth_reg[:, :, 0:1].sigmoid() + 1e-6) - 1.) + depth_geo.unsqueeze(-1) + depth_map) / 3,
                                 depth_reg[:, :, 1:2]], -1))
        angles.append(mlp(sd, f"angle_embed.{lvl}", hs[lvl], 2))
    out = {"pred_logits": classes[-1], "pred_boxes": coords[-1], "pred_3d_dim": dims[-1], "pred_depth": depths[-1],
           "pred_angle": angles[-1], "pred_depth_map_logits": depth_logits}
    out["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b, "pred_3d_dim": c, "pred_angle": d, "pred_depth": e}
                          for a, b, c, d, e in zip(classes[:-1], coords[:-1], dims[:-1], angles[:-1], depths[:-1])]
    return out


# ---------------------------------------------------------------------------------------------------------
# deterministic weights keyed by parameter name (so the reference, this oracle and the product model can be
# filled identically without sharing a 150 MB checkpoint)
# ---------------------------------------------------------------------------------------------------------
def state_dict_spec(cfg=CFG):
    """name -> shape of every parameter/buffer of the reference model (without the decoder alias keys)."""
    spec = {}
    c = cfg["hidden_dim"]
    b = "backbone.0.body."

    def bn(p, n):
        for k in ("weight", "bias", "running_mean", "running_var"):
            spec[f"{p}.{k}"] = (n,)
    spec[b + "conv1.weight"] = (64, 3, 7, 7)
    bn(b + "bn1", 64)
    inplanes = 64
    for name, planes, blocks, stride in RESNET_LAYERS:
        for i in range(blocks):
            p = f"{b}{name}.{i}"
            spec[p + ".conv1.weight"] = (planes, inplanes, 1, 1); bn(p + ".bn1", planes)
            spec[p + ".conv2.weight"] = (planes, planes, 3, 3); bn(p + ".bn2", planes)
            spec[p + ".conv3.weight"] = (planes * 4, planes, 1, 1); bn(p + ".bn3", planes * 4)
            if i == 0:
                spec[p + ".downsample.0.weight"] = (planes * 4, inplanes, 1, 1); bn(p + ".downsample.1", planes * 4)
            inplanes = planes * 4

    def lin(p, o, i):
        spec[p + ".weight"] = (o, i); spec[p + ".bias"] = (o,)

    def norm(p, n=c):
        spec[p + ".weight"] = (n,); spec[p + ".bias"] = (n,)

    def mlp_(p, i, h, o, n):
        dims = [i] + [h] * (n - 1) + [o]
        for k in range(n):
            lin(f"{p}.layers.{k}", dims[k + 1], dims[k])

    def msda(p):
        lin(p + ".sampling_offsets", 256, c); lin(p + ".attention_weights", 128, c)
        lin(p + ".value_proj", c, c); lin(p + ".output_proj", c, c)

    def mha_(p):
        spec[p + ".in_proj_weight"] = (3 * c, c); spec[p + ".in_proj_bias"] = (3 * c,); lin(p + ".out_proj", c, c)
    for l, cin in enumerate((512, 1024, 2048)):
        spec[f"input_proj.{l}.0.weight"] = (c, cin, 1, 1); spec[f"input_proj.{l}.0.bias"] = (c,); norm(f"input_proj.{l}.1")
    spec["input_proj.3.0.weight"] = (c, 2048, 3, 3); spec["input_proj.3.0.bias"] = (c,); norm("input_proj.3.1")
    spec["query_embed.weight"] = (cfg["num_queries"] * cfg["group_num"], 2 * c)
    spec["label_enc.weight"] = (cfg["num_classes"] + 1, c - 1)
    for l in range(cfg["dec_layers"]):
        lin(f"class_embed.{l}", cfg["num_classes"], c)
        mlp_(f"bbox_embed.{l}", c, c, 6, 3); mlp_(f"dim_embed_3d.{l}", c, c, 3, 2)
        mlp_(f"angle_embed.{l}", c, c, 24, 2); mlp_(f"depth_embed.{l}", c, c, 2, 2)
    d = "depth_predictor."
    spec[d + "depth_bin_values"] = (cfg["num_depth_bins"] + 1,)
    for nme, k in (("downsample", 3), ("proj", 1), ("upsample", 1)):
        spec[f"{d}{nme}.0.weight"] = (c, c, k, k); spec[f"{d}{nme}.0.bias"] = (c,); norm(f"{d}{nme}.1")
    for i in (0, 3):
        spec[f"{d}depth_head.{i}.weight"] = (c, c, 3, 3); spec[f"{d}depth_head.{i}.bias"] = (c,); norm(f"{d}depth_head.{i + 1}")
    spec[d + "depth_classifier.weight"] = (cfg["num_depth_bins"] + 1, c, 1, 1); spec[d + "depth_classifier.bias"] = (cfg["num_depth_bins"] + 1,)
    e = d + "depth_encoder.layers.0"
    mha_(e + ".self_attn"); lin(e + ".linear1", 256, c); lin(e + ".linear2", c, 256); norm(e + ".norm1"); norm(e + ".norm2")
    spec[d + "depth_pos_embed.weight"] = (int(cfg["depth_max"]) + 1, 256)
    t = "depthaware_transformer."
    spec[t + "level_embed"] = (cfg["num_feature_levels"], c)
    lin(t + "reference_points", 2, c)
    ff = cfg["dim_feedforward"]
    for l in range(cfg["enc_layers"]):
        p = f"{t}encoder.layers.{l}"
        msda(p + ".self_attn"); norm(p + ".norm1"); lin(p + ".linear1", ff, c); lin(p + ".linear2", c, ff); norm(p + ".norm2")
    for l in range(cfg["dec_layers"]):
        p = f"{t}decoder.layers.{l}"
        msda(p + ".cross_attn"); norm(p + ".norm1"); mha_(p + ".cross_attn_depth"); norm(p + ".norm_depth")
        mha_(p + ".self_attn"); norm(p + ".norm2"); lin(p + ".linear1", ff, c); lin(p + ".linear2", c, ff); norm(p + ".norm3")
        for nme in ("sa_qcontent_proj", "sa_qpos_proj", "sa_kcontent_proj", "sa_kpos_proj", "sa_v_proj"):
            lin(f"{p}.{nme}", c, c)
    mlp_(t + "decoder.query_scale", c, c, c, 2)
    mlp_(t + "decoder.ref_point_head", c, c, 2, 2)
    return spec


ALIASES = {"depthaware_transformer.decoder.bbox_embed.": "bbox_embed.", "depthaware_transformer.decoder.dim_embed.": "dim_embed_3d."}


def _seed_of(name):
    h = 1469598103934665603
    for ch in name.encode():
        h = ((h ^ ch) * 1099511628211) % (1 << 63)
    return h % (2 ** 31 - 1)


def deterministic_state_dict(cfg=CFG, base_seed=0, dtype=torch.float32):
    """Well-conditioned pseudo-random weights keyed by NAME (independent of construction order)."""
    sd = {}
    for name, shape in state_dict_spec(cfg).items():
        g = torch.Generator().manual_seed(_seed_of(name) + base_seed)
        leaf = name.rsplit(".", 1)[-1]
        if name.endswith("depth_bin_values"):
            nb, dmin, dmax = cfg["num_depth_bins"], cfg["depth_min"], cfg["depth_max"]
            bin_size = 2 * (dmax - dmin) / (nb * (1 + nb))
            idx = torch.linspace(0, nb - 1, nb)
            t = torch.cat([(idx + 0.5).pow(2) * bin_size / 2 - bin_size / 8 + dmin, torch.tensor([dmax])])
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif leaf == "running_mean":
            t = torch.randn(shape, generator=g) * 0.1
        elif ".bn" in name or "downsample.1" in name and "backbone" in name:
            t = (torch.rand(shape, generator=g) * 0.5 + 0.75) if leaf == "weight" else torch.randn(shape, generator=g) * 0.1
        elif len(shape) == 1:
            is_norm_w = leaf == "weight"
            t = (torch.rand(shape, generator=g) * 0.5 + 0.75) if is_norm_w else torch.randn(shape, generator=g) * 0.05
        elif "embed.weight" in name or name.endswith("level_embed") or name.endswith("label_enc.weight"):
            t = torch.randn(shape, generator=g) * (1.0 if "depth_pos" not in name else 0.5)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * (1.6 / fan_in) ** 0.5
            if name.endswith("sampling_offsets.weight"):
                t = t * 0.3
        sd[name] = t.to(dtype)
    # keep sampling offsets spread like the reference init (ms_deform_attn.py:106-114) so samples stay local
    for name in list(sd):
        if name.endswith("sampling_offsets.bias"):
            thetas = torch.arange(8, dtype=torch.float32) * (2.0 * math.pi / 8)
            grid = torch.stack([thetas.cos(), thetas.sin()], -1)
            grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(8, 1, 1, 2).repeat(1, 4, 4, 1)
            for i in range(4):
                grid[:, :, i, :] *= i + 1
            sd[name] = grid.view(-1).to(dtype)
    return sd


def with_aliases(sd):
    out = dict(sd)
    for k, v in sd.items():
        for alias, real in ALIASES.items():
            if k.startswith(real):
                out[alias + k[len(real):]] = v
    return out


def synthetic_inputs(B, seed=0, H=384, W=1280, device="cpu"):
    """N(0,1) images, KITTI P2 calib, img_sizes (1242, 375)."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(B, 3, H, W, generator=g)
    calibs = torch.zeros(B, 3, 4)
    calibs[:, 0, 0] = calibs[:, 1, 1] = 721.5377
    calibs[:, 0, 2] = 609.5593; calibs[:, 1, 2] = 172.854; calibs[:, 0, 3] = 44.85728
    img_sizes = torch.tensor([[1242., 375.]]).repeat(B, 1)
    return images.to(device), calibs.to(device), img_sizes.to(device)


def surrogate_loss(out):
    """Device-agnostic surrogate L = sum_k mean(out[k]^2) over all tensor outputs incl. aux."""
    loss = 0.0
    for k, v in out.items():
        if k == "aux_outputs":
            for aux in v:
                for t in aux.values():
                    loss = loss + (t ** 2).mean()
        else:
            loss = loss + (v ** 2).mean()
    return loss


def bench_reference_model(args):
    """`bench.py --impl reference --workload model`: train-shape forward+backward of this CPU port on a bounded sample."""
    Bs = getattr(args, "cpu_batch", 2)
    sd = deterministic_state_dict()
    # Thread count: SURVEY.md 8(d) asks for all host cores, but torch's CPU kernels stop scaling (and regress) well before
    # 128 logical cores; sweep once on a small forward pass (half resolution) and time the real sample at the best count.
    if os.environ.get("MDB_CPU_THREADS"):
        threads, sweep = min(os.cpu_count(), int(os.environ["MDB_CPU_THREADS"])), None
    else:
        cands = sorted({c for c in (16, 32, 64, os.cpu_count()) if c <= os.cpu_count()})
        im_s, ca_s, sz_s = synthetic_inputs(1, 0, H=192, W=640)
        sweep = {}
        for c in cands:
            torch.set_num_threads(c)
            with torch.no_grad():
                forward(sd, im_s, ca_s, sz_s, training=True)          # warm-up at this count
                t0 = time.time()
                forward(sd, im_s, ca_s, sz_s, training=True)
                sweep[c] = round(time.time() - t0, 3)
        threads = min(sweep, key=sweep.get)
    torch.set_num_threads(threads)
    trainable = [k for k in sd if sd[k].dtype.is_floating_point and "running_" not in k and "depth_bin_values" not in k
                 and not (k.startswith("backbone.0.body.") and not any(s in k for s in ("layer2", "layer3", "layer4")))
                 and ".bn" not in k and "downsample.1" not in k]
    for k in trainable:
        sd[k].requires_grad_(True)
    images, calibs, sizes = synthetic_inputs(Bs, 0)

    def step():
        for k in trainable:
            sd[k].grad = None
        out = forward(sd, images, calibs, sizes, training=True)
        surrogate_loss(out).backward()
    for _ in range(min(args.warmup, 1)):
        step()
    # bounded: at most `steps` passes and at most ~budget seconds of CPU work (always at least one pass)
    budget = float(os.environ.get("MDB_CPU_BUDGET_S", "150"))
    t0 = time.time()
    done = 0
    while done < args.steps and (done == 0 or time.time() - t0 < budget):
        step()
        done += 1
    dt = (time.time() - t0) * args.steps / done      # scaled to the requested step count (ms_per_step stays per pass)
    cfg = {"workload": "full MonoDETR fwd+bwd (train shapes: 550 queries, group self-attn), ResNet-50, 1280x384 synthetic, "
                       "CPU port of the reference path, surrogate loss"}
    sample = (f"B={Bs} per step (bounded sample of the batch-8 workload), {done} of {args.steps} steps actually run within the "
              f"{budget:.0f} s CPU budget; threads = {threads} of {os.cpu_count()} logical cores"
              + (f" (fastest of a one-off sweep, seconds per half-resolution forward: {sweep})" if sweep else " (MDB_CPU_THREADS)"))
    return Bs * args.steps / dt, dt, sample, threads, cfg
