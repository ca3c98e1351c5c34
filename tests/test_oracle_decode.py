"""The numpy restatement of the post-process (oracle/decode.py) against vectors produced by the unmodified reference
functions (tests/golden/decode.npz, tools/gen_golden_decode.py)."""
import os

import numpy as np
import pytest

from oracle import decode as od

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "decode.npz"))
CASES = ["eval", "train_queries", "low_threshold"]


def angle_close(a, b, tol):
    return np.all(np.abs(np.angle(np.exp(1j * (np.asarray(a) - np.asarray(b))))) <= tol)


@pytest.mark.parametrize("name", CASES)
def test_extract_matches_reference(name):
    seed, B, Q, topk, thr = GOLD[f"{name}.cfg"]
    h = od.synthetic_heads(int(seed), int(B), int(Q))
    dets = od.extract_dets(h["logits"], h["boxes"], h["dim3"], h["depth"], h["angle"], int(topk))
    ref = GOLD[f"{name}.dets"]
    assert np.array_equal(dets[..., 0], ref[..., 0])                          # labels: same candidates in the same order
    np.testing.assert_allclose(dets, ref, rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("name", CASES)
def test_decode_matches_reference(name):
    seed, B, Q, topk, thr = GOLD[f"{name}.cfg"]
    h = od.synthetic_heads(int(seed), int(B), int(Q))
    res = od.decode_dets(GOLD[f"{name}.dets"], h["img_size"], h["P2"], h["mean_size"], float(thr))
    rows, count = GOLD[f"{name}.rows"], GOLD[f"{name}.count"]
    assert [len(r) for r in res] == count.tolist()
    assert count.sum() > 0
    for i, r in enumerate(res):
        if not r:
            continue
        r = np.array(r)
        g = rows[i, :len(r)]
        lin = [0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13]
        np.testing.assert_allclose(r[:, lin], g[:, lin], rtol=1e-5, atol=1e-5)   # the reference mixes fp32 and fp64 scalars
        assert angle_close(r[:, 1], g[:, 1], 1e-5) and angle_close(r[:, 12], g[:, 12], 1e-5)
