"""Device post-process (csrc/decode.cu through monodetr_b200.decode) against the oracle and the reference's golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import decode as od

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "decode.npz"))
CASES = ["eval", "train_queries", "low_threshold"]


def _outputs(h, dev="cuda"):
    t = lambda k: torch.from_numpy(h[k]).to(dev)  # noqa: E731
    return {"pred_logits": t("logits"), "pred_boxes": t("boxes"), "pred_3d_dim": t("dim3"), "pred_depth": t("depth"),
            "pred_angle": t("angle")}


def _angle_err(a, b):
    return np.abs(np.angle(np.exp(1j * (np.asarray(a, np.float64) - np.asarray(b, np.float64)))))


@pytest.mark.parametrize("name", CASES)
def test_extract_against_reference_golden(name):
    from monodetr_b200 import decode
    seed, B, Q, topk, thr = GOLD[f"{name}.cfg"]
    h = od.synthetic_heads(int(seed), int(B), int(Q))
    dets = decode.extract_dets_from_outputs(_outputs(h), K=50, topk=int(topk)).cpu().numpy()
    ref = GOLD[f"{name}.dets"]
    assert np.array_equal(dets[..., 0], ref[..., 0])
    np.testing.assert_allclose(dets, ref, rtol=3e-6, atol=1e-7)               # expf on the device vs torch's CPU sigmoid / exp


@pytest.mark.parametrize("name", CASES)
def test_decode_against_reference_golden(name):
    from monodetr_b200 import decode
    seed, B, Q, topk, thr = GOLD[f"{name}.cfg"]
    B = int(B)
    h = od.synthetic_heads(int(seed), B, int(Q))
    dets = torch.from_numpy(GOLD[f"{name}.dets"]).cuda()
    rows, count = decode.decode_detections_device(dets, h["img_size"], h["P2"], h["mean_size"], float(thr))
    rows, count = rows.cpu().numpy(), count.cpu().numpy()
    g_rows, g_count = GOLD[f"{name}.rows"], GOLD[f"{name}.count"]
    assert count.tolist() == g_count.tolist()
    lin = [0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13]
    for i in range(B):
        n = int(count[i])
        # fp32 on the device, mixed fp32 / fp64 in the reference: pixel coordinates reach ~1e3, so 1e-3 absolute is 1 ulp-ish
        np.testing.assert_allclose(rows[i, :n][:, lin], g_rows[i, :n][:, lin], rtol=2e-5, atol=2e-3)
        assert _angle_err(rows[i, :n, 1], g_rows[i, :n, 1]).max(initial=0) < 1e-5
        assert _angle_err(rows[i, :n, 12], g_rows[i, :n, 12]).max(initial=0) < 1e-5
        assert not rows[i, n:].any()                                          # rows beyond the count are zero-filled


def test_reference_signature_end_to_end_against_oracle():
    """extract -> decode through the reference-shaped API, calibs as objects with .P2, random batch sizes / thresholds."""
    from monodetr_b200 import decode

    class Calib:
        def __init__(self, P2):
            self.P2 = P2

    for seed, B, Q, topk, thr in [(1, 1, 50, 50, 0.2), (2, 8, 50, 50, 0.05), (3, 3, 550, 100, 0.5), (4, 2, 7, 21, 0.0)]:
        h = od.synthetic_heads(seed, B, Q)
        dets = decode.extract_dets_from_outputs(_outputs(h), topk=topk)
        info = {"img_id": np.arange(100, 100 + B), "img_size": h["img_size"]}
        res = decode.decode_detections(dets, info, [Calib(h["P2"][i]) for i in range(B)], h["mean_size"], thr)
        want = od.decode_dets(od.extract_dets(h["logits"], h["boxes"], h["dim3"], h["depth"], h["angle"], topk), h["img_size"], h["P2"],
                              h["mean_size"], thr)
        assert sorted(res) == list(range(100, 100 + B))
        for i in range(B):
            got = res[100 + i]
            assert len(got) == len(want[i])
            if not got:
                continue
            assert isinstance(got[0][0], int)
            a, b = np.array(got, np.float64), np.array(want[i], np.float64)
            lin = [0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13]
            np.testing.assert_allclose(a[:, lin], b[:, lin], rtol=2e-5, atol=2e-3)
            assert _angle_err(a[:, 1], b[:, 1]).max() < 1e-5 and _angle_err(a[:, 12], b[:, 12]).max() < 1e-5


def test_ties_nan_and_limits():
    from monodetr_b200 import decode
    h = od.synthetic_heads(5, 2, 50)
    h["logits"][:] = 0.25                                                     # every candidate ties: flat index order
    h["logits"][1, 3, 1] = np.nan                                             # a NaN logit sorts last
    dets = decode.extract_dets_from_outputs(_outputs(h), topk=50).cpu().numpy()
    assert np.array_equal(dets[0, :, 0], np.arange(50) % 3)
    np.testing.assert_allclose(dets[0, :, 34], h["boxes"][0, np.arange(50) // 3, 0])
    assert np.isfinite(dets[1, :, 1]).all()
    with pytest.raises(RuntimeError):
        decode.extract_dets_from_outputs({k: v.cpu() for k, v in _outputs(h).items()})
    big = od.synthetic_heads(6, 1, 2000)                                      # 6000 candidates > the 4096 the kernel sorts
    with pytest.raises(RuntimeError):
        decode.extract_dets_from_outputs(_outputs(big))


def test_model_outputs_to_detections():
    """The eval path a tester runs: MonoDETR.forward -> extract -> decode, all on the device, one copy at the end."""
    from monodetr_b200 import build_monodetr, decode
    from monodetr_b200.monodetr import DEFAULT_MODEL_CFG
    torch.manual_seed(0)
    model = build_monodetr(dict(DEFAULT_MODEL_CFG))[0].cuda().eval()
    B = 2
    imgs = torch.randn(B, 3, 192, 640, device="cuda")
    h = od.synthetic_heads(7, B, 50)
    calibs = torch.from_numpy(h["P2"]).cuda()
    sizes = torch.from_numpy(h["img_size"]).cuda()
    with torch.no_grad():
        out = model(imgs, calibs, None, sizes)
    dets = decode.extract_dets_from_outputs(out, topk=50)
    res = decode.decode_detections(dets, {"img_id": [0, 1], "img_size": sizes}, calibs, h["mean_size"], 0.0)
    o = {k: out[k].float().cpu().numpy() for k in ("pred_logits", "pred_boxes", "pred_3d_dim", "pred_depth", "pred_angle")}
    want = od.decode_dets(od.extract_dets(o["pred_logits"], o["pred_boxes"], o["pred_3d_dim"], o["pred_depth"], o["pred_angle"], 50),
                          h["img_size"], h["P2"], h["mean_size"], 0.0)
    for i in range(B):
        assert len(res[i]) == len(want[i]) == 50
        a, b = np.array(res[i], np.float64), np.array(want[i], np.float64)
        assert np.array_equal(a[:, 0], b[:, 0])
        np.testing.assert_allclose(a[:, [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13]], b[:, [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13]], rtol=2e-5, atol=2e-3)
