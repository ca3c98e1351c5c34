"""oracle/preprocess.py and the host-side get_affine_transform against Pillow / cv2 outputs produced with the reference's own calls
(tests/golden/preprocess.npz, tools/gen_golden_preprocess.py)."""
import os

import numpy as np

from oracle import preprocess as op

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz"))


def test_warp_and_normalize_are_bit_identical_to_pillow_and_numpy():
    sizes = [tuple(s) for s in GOLD["sizes"]]
    imgs = op.synthetic_images(7, sizes)
    res = tuple(int(v) for v in GOLD["resolution"])
    for i, im in enumerate(imgs):
        src = im[:, ::-1] if bool(GOLD[f"{i}.flip"]) else im
        u8 = op.warp_affine_bilinear(src, GOLD[f"{i}.trans_inv"].reshape(-1), res)
        assert np.array_equal(u8, GOLD[f"{i}.u8"]), i
        if f"{i}.normalized" in GOLD.files:
            assert np.array_equal(op.normalize(u8), GOLD[f"{i}.normalized"])
    assert (GOLD["3.u8"] == 0).all(axis=-1).any()                     # the shifted crops really exercise the zero fill


def test_get_affine_transform_matches_the_reference():
    from monodetr_b200.preprocess import get_affine_transform
    res = GOLD["resolution"]
    for i in range(len(GOLD["sizes"])):
        trans, trans_inv = get_affine_transform(GOLD[f"{i}.center"], GOLD[f"{i}.crop_size"], 0, res, inv=1)
        np.testing.assert_allclose(trans, GOLD[f"{i}.trans"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(trans_inv, GOLD[f"{i}.trans_inv"], rtol=0, atol=1e-9)
