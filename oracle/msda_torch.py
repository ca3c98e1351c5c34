"""TEST / BASELINE INFRASTRUCTURE: torch (CPU-capable, multi-threaded) restatement of the reference's only
CPU implementation of the MSDeformAttn core, `ms_deform_attn_core_pytorch`
(lib/models/monodetr/ops/functions/ms_deform_attn_func.py:41-61): per level, F.grid_sample(bilinear,
zeros, align_corners=False) of the (N*M, D, H, W) value map on grid 2*loc-1, then the attention-weighted
sum over (level, point).  Differentiable through autograd (that is how the reference's CPU backward runs).
Pinned against the reference in tests/test_oracle_msda.py::test_torch_port_matches_reference.
"""
import torch
import torch.nn.functional as F


def msda_core_torch(value, spatial_shapes, sampling_locations, attention_weights):
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes) else spatial_shapes)]
    grids = sampling_locations * 2 - 1
    start = 0
    sampled = []
    for lvl, (H, W) in enumerate(shapes):
        v = value[:, start:start + H * W].permute(0, 2, 3, 1).reshape(N * M, D, H, W)
        start += H * W
        g = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(N * M, Lq, P, 2)
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    stacked = torch.stack(sampled, dim=3).reshape(N * M, D, Lq, L * P)          # (N*M, D, Lq, L*P)
    w = attention_weights.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L * P)
    out = (stacked * w).sum(-1).view(N, M * D, Lq)
    return out.transpose(1, 2).contiguous()
