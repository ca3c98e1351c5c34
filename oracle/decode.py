"""CPU restatement (numpy) of the reference's inference post-process -- TEST INFRASTRUCTURE ONLY (imported by tests/ and
tools/gen_golden_decode.py; the product path is monodetr_b200/csrc/decode.cu).

  extract_dets   lib/helpers/decode_helper.py:57-110   (extract_dets_from_outputs)
  decode_dets    lib/helpers/decode_helper.py:8-54     (decode_detections) with the camera arithmetic of
                 lib/datasets/kitti/kitti_utils.py:150-155 (intrinsics), :207-208 (img_to_rect), :277-282 (alpha2ry) and
                 lib/datasets/utils.py:19-26 (class2angle), decode_helper.py:174-178 (get_heading_angle)

Pinned against the unmodified reference functions by tests/golden/decode.npz (tools/gen_golden_decode.py imports them from
/root/reference) -- tests/test_oracle_decode.py.
"""
import numpy as np

NUM_HEADING_BIN = 12      # lib/datasets/utils.py:7


def extract_dets(logits, boxes, dim3, depth, angle, topk=50):
    """(B,Q,C), (B,Q,6), (B,Q,3), (B,Q,2), (B,Q,24) -> (B, topk, 37) float32; ties keep the lower flat index first."""
    logits = np.asarray(logits, np.float32)
    B, Q, C = logits.shape
    prob = (1.0 / (1.0 + np.exp(-logits.astype(np.float64)))).astype(np.float32).reshape(B, Q * C)
    order = np.argsort(-logits.reshape(B, Q * C), axis=1, kind="stable")[:, :topk]      # decode_helper.py:66
    out = np.zeros((B, topk, 37), np.float32)
    for b in range(B):
        flat = order[b]
        q, lab = flat // C, flat % C                                                        # :71-73
        bx = np.asarray(boxes[b], np.float32)[q]
        x0, y0 = bx[:, 0] - bx[:, 2], bx[:, 1] - bx[:, 4]                                   # utils/box_ops.py:20-24
        x1, y1 = bx[:, 0] + bx[:, 3], bx[:, 1] + bx[:, 5]
        out[b, :, 0] = lab
        out[b, :, 1] = prob[b, flat]
        out[b, :, 2] = (x0 + x1) / 2                                                       # utils/box_ops.py:27-31
        out[b, :, 3] = (y0 + y1) / 2
        out[b, :, 4] = x1 - x0
        out[b, :, 5] = y1 - y0
        out[b, :, 6] = np.asarray(depth[b], np.float32)[q, 0]
        out[b, :, 7:31] = np.asarray(angle[b], np.float32)[q]
        out[b, :, 31:34] = np.asarray(dim3[b], np.float32)[q]
        out[b, :, 34] = bx[:, 0]
        out[b, :, 35] = bx[:, 1]
        out[b, :, 36] = np.exp(-np.asarray(depth[b], np.float32)[q, 1])                      # :79
    return out


def decode_dets(dets, img_size, P2, cls_mean_size, threshold):
    """dets (B,K,37), img_size (B,2) [W,H], P2 (B,3,4), cls_mean_size (C,3) -> list (per image) of rows
    [cls, alpha, x0, y0, x1, y1, h, w, l, X, Y, Z, ry, score] in float64."""
    dets = np.asarray(dets, np.float64)
    results = []
    for i in range(dets.shape[0]):
        P = np.asarray(P2[i], np.float64)
        cu, cv, fu, fv = P[0, 2], P[1, 2], P[0, 0], P[1, 1]
        tx, ty = P[0, 3] / (-fu), P[1, 3] / (-fv)
        W, H = float(img_size[i][0]), float(img_size[i][1])
        rows = []
        for j in range(dets.shape[1]):
            d = dets[i, j]
            cls_id, score = int(d[0]), d[1]
            if score < threshold:
                continue
            x, y, w, h = d[2] * W, d[3] * H, d[4] * W, d[5] * H
            depth = d[6]
            dims = d[31:34] + np.asarray(cls_mean_size, np.float64)[cls_id]
            x3d, y3d = d[34] * W, d[35] * H
            loc = [((x3d - cu) * depth) / fu + tx, ((y3d - cv) * depth) / fv + ty + dims[0] / 2, depth]
            cls = int(np.argmax(d[7:7 + NUM_HEADING_BIN]))
            alpha = cls * (2 * np.pi / NUM_HEADING_BIN) + d[7 + NUM_HEADING_BIN + cls]
            if alpha > np.pi:
                alpha -= 2 * np.pi
            ry = alpha + np.arctan2(x - cu, fu)
            if ry > np.pi:
                ry -= 2 * np.pi
            if ry < -np.pi:
                ry += 2 * np.pi
            rows.append([cls_id, alpha, x - w / 2, y - h / 2, x + w / 2, y + h / 2] + dims.tolist() + loc + [ry, score * d[36]])
        results.append(rows)
    return results


def synthetic_heads(seed, B, Q, C=3):
    """Seeded head outputs shaped like MonoDETR's (monodetr.py:270-283) + KITTI-like calibration, for the parity tests."""
    g = np.random.default_rng(seed)
    logits = g.normal(-2.0, 2.5, (B, Q, C)).astype(np.float32)
    cxcy = g.uniform(0.1, 0.9, (B, Q, 2))
    lrtb = g.uniform(0.01, 0.2, (B, Q, 4))
    boxes = np.concatenate([cxcy, lrtb], -1).astype(np.float32)
    dim3 = g.normal(0, 0.3, (B, Q, 3)).astype(np.float32)
    depth = np.stack([g.uniform(3, 60, (B, Q)), g.normal(0, 1, (B, Q))], -1).astype(np.float32)
    angle = np.concatenate([g.normal(0, 2, (B, Q, 12)), g.uniform(-0.3, 0.3, (B, Q, 12))], -1).astype(np.float32)
    img_size = np.tile(np.array([[1242.0, 375.0]], np.float32), (B, 1)) + g.integers(-20, 20, (B, 2)).astype(np.float32)
    P2 = np.tile(np.array([[721.5377, 0, 609.5593, 44.85728], [0, 721.5377, 172.854, 0.2163791], [0, 0, 1, 0.002745884]],
                          np.float32), (B, 1, 1))
    P2[:, 0, 2] += g.normal(0, 3, B).astype(np.float32)
    P2[:, 0, 3] += g.normal(0, 1, B).astype(np.float32)
    mean_size = np.array([[1.76255119, 0.66068622, 0.84422524], [1.52563191462, 1.62856739989, 3.88311640418],
                          [1.73698127, 0.59706367, 1.76282397]], np.float32)          # kitti_dataset.py:75-77
    return dict(logits=logits, boxes=boxes, dim3=dim3, depth=depth, angle=angle, img_size=img_size, P2=P2, mean_size=mean_size)
