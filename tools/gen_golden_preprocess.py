"""Golden vectors for the input pipeline, produced with the reference's own calls: get_affine_transform
(lib/datasets/kitti/kitti_utils.py:347-381, cv2) and PIL's Image.transform / transpose + the numpy normalisation as written at
lib/datasets/kitti/kitti_dataset.py:140-161.  Run in the build container:  python tools/gen_golden_preprocess.py"""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from oracle.preprocess import MEAN, STD, synthetic_images  # noqa: E402
from lib.datasets.kitti.kitti_utils import get_affine_transform  # noqa: E402

resolution = np.array([320, 96])
sizes = [(311, 94), (306, 93), (311, 94), (200, 120), (311, 94)]
imgs = synthetic_images(7, sizes)
rng = np.random.default_rng(8)
out = {"resolution": resolution, "sizes": np.array(sizes)}
for i, im in enumerate(imgs):
    img = Image.fromarray(im)
    img_size = np.array(img.size)
    center = np.array(img_size) / 2
    crop_size, flip = img_size, False
    if i in (1, 4):
        flip = True
        img = img.transpose(Image.FLIP_LEFT_RIGHT)
    if i >= 2:                                           # the random crop of kitti_dataset.py:144-150 (scale 0.4, shift 0.1)
        crop_scale = np.clip(rng.standard_normal() * 0.4 + 1, 0.6, 1.4)
        crop_size = img_size * crop_scale
        center[0] += img_size[0] * np.clip(rng.standard_normal() * 0.1, -0.2, 0.2)
        center[1] += img_size[1] * np.clip(rng.standard_normal() * 0.1, -0.2, 0.2)
    trans, trans_inv = get_affine_transform(center, crop_size, 0, resolution, inv=1)
    warped = img.transform(tuple(resolution.tolist()), method=Image.AFFINE, data=tuple(trans_inv.reshape(-1).tolist()), resample=Image.BILINEAR)
    u8 = np.array(warped)
    x = u8.astype(np.float32) / 255.0
    x = ((x - MEAN) / STD).transpose(2, 0, 1)
    out[f"{i}.center"], out[f"{i}.crop_size"], out[f"{i}.flip"] = center, np.asarray(crop_size, np.float64), np.array(flip)
    out[f"{i}.trans"], out[f"{i}.trans_inv"], out[f"{i}.u8"] = trans, trans_inv, u8
    if i in (0, 4):
        out[f"{i}.normalized"] = x
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess.npz"), **out)
print({k: v.shape for k, v in out.items() if k.endswith("u8")})
