"""Device input pipeline (csrc/preprocess.cu through monodetr_b200.preprocess) against Pillow's golden outputs and the numpy oracle:
bit-exact (8-bit interpolation result and fp32 normalisation)."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as op

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz"))


def test_against_pillow_golden_bit_exact():
    from monodetr_b200.preprocess import ImageBatchPreprocessor
    sizes = [tuple(s) for s in GOLD["sizes"]]
    imgs = op.synthetic_images(7, sizes)
    B = len(imgs)
    pre = ImageBatchPreprocessor(resolution=tuple(int(v) for v in GOLD["resolution"]))
    tinv = np.stack([GOLD[f"{i}.trans_inv"] for i in range(B)])
    flip = [bool(GOLD[f"{i}.flip"]) for i in range(B)]
    out = pre([torch.from_numpy(im) for im in imgs], tinv, flip).cpu().numpy()
    for i in range(B):
        assert np.array_equal(out[i], op.normalize(GOLD[f"{i}.u8"])), i
    assert np.array_equal(out[0], GOLD["0.normalized"]) and np.array_equal(out[4], GOLD["4.normalized"])


def test_full_resolution_ragged_batch_against_oracle():
    """KITTI-sized ragged batch -> 1280x384, random crop / shift / flip as kitti_dataset.py:140-150; CUDA and CPU source tensors."""
    from monodetr_b200.preprocess import ImageBatchPreprocessor, get_affine_transform
    sizes = [(1242, 375), (1224, 370), (1238, 374), (1241, 376)]
    imgs = op.synthetic_images(9, sizes)
    rng = np.random.default_rng(10)
    res = np.array([1280, 384])
    tinv, flip = [], []
    for (W, H) in sizes:
        size = np.array([W, H], np.float64)
        center = size / 2
        scale = np.clip(rng.standard_normal() * 0.4 + 1, 0.6, 1.4)
        center += size * np.clip(rng.standard_normal(2) * 0.1, -0.2, 0.2)
        tinv.append(get_affine_transform(center, size * scale, 0, res, inv=1)[1])
        flip.append(bool(rng.integers(0, 2)))
    pre = ImageBatchPreprocessor(resolution=(1280, 384))
    srcs = [torch.from_numpy(im).cuda() if i % 2 else torch.from_numpy(im) for i, im in enumerate(imgs)]
    out = pre(srcs, np.stack(tinv), flip)
    assert out.shape == (4, 3, 384, 1280) and out.dtype == torch.float32
    out = out.cpu().numpy()
    for i, im in enumerate(imgs):
        want = op.preprocess(im, tinv[i].reshape(-1), (1280, 384), flip[i])
        assert np.array_equal(out[i], want), (i, np.abs(out[i] - want).max(), (out[i] != want).mean())


def test_errors():
    from monodetr_b200.preprocess import ImageBatchPreprocessor
    pre = ImageBatchPreprocessor(resolution=(64, 32))
    with pytest.raises(ValueError):
        pre([torch.zeros(8, 8, 3)], np.eye(2, 3)[None])                 # float image
    with pytest.raises(RuntimeError):
        ImageBatchPreprocessor(device="cpu")([torch.zeros(8, 8, 3, dtype=torch.uint8)], np.eye(2, 3)[None])
