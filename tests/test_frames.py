"""CPU: frame-folder loader (lib/datasets/frames.py) against the reference's own image pipeline
(dataset.py:238-262: torchvision Resize(NEAREST) + CenterCrop on PIL images, Normalize) on synthetic ScanNet folders."""
import math
import os

import numpy as np
import pytest
import torch
from PIL import Image
from torchvision import transforms

from lib.datasets import frames as F
from lib.utils.config import cfg


def _ref_resize_crop(image, new_dims):  # the reference's calls, dataset.py:238-246
    dims = [image.shape[1], image.shape[0]]
    if dims == list(new_dims):
        return image
    rw = int(math.floor(new_dims[1] * float(dims[0]) / float(dims[1])))
    im = transforms.Resize([new_dims[1], rw], interpolation=Image.NEAREST)(Image.fromarray(image))
    return np.array(transforms.CenterCrop([new_dims[1], new_dims[0]])(im))


@pytest.mark.parametrize("hw,new", [((480, 640), [41, 32]), ((968, 1296), [328, 256]), ((240, 320), [328, 256]), ((32, 41), [41, 32]),
                                    ((481, 643), [41, 32]), ((256, 330), [328, 256]), ((97, 131), [41, 32])])
def test_resize_crop_matches_torchvision_nearest(hw, new):
    rng = np.random.default_rng(hw[0])
    depth = rng.integers(0, 6000, hw).astype(np.uint16)
    # the reference reads depth PNGs with scipy.misc.imread, which hands 16-bit images over as int32 (PIL mode "I")
    assert np.array_equal(F.resize_crop_image(depth, new), _ref_resize_crop(depth.astype(np.int32), new))
    colour = rng.integers(0, 256, hw + (3,)).astype(np.uint8)
    assert np.array_equal(F.resize_crop_image(colour, new), _ref_resize_crop(colour, new))


def _make_scene(root, name, n, rng):
    d = os.path.join(root, name)
    for sub in ("depth", "pose", "color2"):
        os.makedirs(os.path.join(d, sub))
    want = []
    for fid in range(0, 20 * n, 20):
        depth = rng.integers(300, 5000, (480, 640)).astype(np.uint16)
        Image.fromarray(depth).save(os.path.join(d, "depth", f"{fid}.png"))
        col = rng.integers(0, 256, (484, 648, 3)).astype(np.uint8)
        Image.fromarray(col).save(os.path.join(d, "color2", f"{fid}.png"))
        pose = np.eye(4, dtype=np.float32)
        pose[:3, 3] = rng.uniform(-2, 2, 3)
        np.savetxt(os.path.join(d, "pose", f"{fid}.txt"), pose, fmt="%.6f")
        want.append((fid, depth, col, np.loadtxt(os.path.join(d, "pose", f"{fid}.txt")).astype(np.float32)))
    w2g = np.eye(4, dtype=np.float32) * 21.333
    w2g[3, 3] = 1
    w2g[:3, 3] = [60.0, 40.0, 55.5]
    np.savetxt(os.path.join(d, "world2grid.txt"), w2g, fmt="%.4f")
    return want, w2g


def test_frame_folders_chunk_and_scene_mode(tmp_path):
    rng = np.random.default_rng(0)
    base = str(tmp_path / "frames_square")
    want, w2g = _make_scene(base, "scene0011_00", 3, rng)
    old = cfg.IMAGE_EXT
    cfg.IMAGE_EXT = ".png"  # lossless synthetic frames
    try:
        chunk_w2g = np.eye(4, dtype=np.float32) * 2
        prov = F.FrameFolders(base, mode="chunk")
        v = prov("/data/chunks/scene0011_00__0__3.chunk", [20, 0], chunk_w2g)
        assert v["frameids"] == [20, 0] and np.array_equal(v["world2grid"], chunk_w2g)
        assert v["depths"].shape == (2, 32, 41) and v["images"].shape == (2, 3, 256, 328) and v["poses"].shape == (2, 4, 4)
        for k, fid in enumerate((20, 0)):
            _, depth, col, pose = next(w for w in want if w[0] == fid)
            assert np.array_equal(v["depths"][k], _ref_resize_crop(depth.astype(np.int32), [41, 32]).astype(np.float32) / 1000.0)
            ref = _ref_resize_crop(col, [328, 256])
            ref = transforms.Normalize(mean=cfg.COLOR_MEAN, std=cfg.COLOR_STD)(torch.Tensor(np.transpose(ref, [2, 0, 1]).astype(np.float32) / 255.0))
            assert np.allclose(v["images"][k], ref.numpy(), atol=2e-6, rtol=1e-6)
            assert np.array_equal(v["poses"][k], pose)
        # whole-scene mode: every depth frame of the folder, folder world2grid minus the volume padding
        feats = lambda x: np.zeros((x.shape[0], 128, 32, 41), np.float32)  # stand-in for an ENet encoder
        v = F.FrameFolders(base, mode="scene", features=feats)("/data/scenes/scene0011_00__0.scene", None, None)
        assert v["frameids"] == ["0", "20", "40"] and v["images"].shape == (3, 128, 32, 41)
        assert np.allclose(v["world2grid"][:3, 3], w2g[:3, 3] - np.array([10, 16, 10]), atol=1e-3)
    finally:
        cfg.IMAGE_EXT = old


def test_scene_name_rules():
    assert F.scene_name_of("/x/scene0011_00__0__3.chunk", "/d/frames_square", "chunk") == "scene0011_00"
    assert F.scene_name_of("/x/abc123_7.chunk", "/d/augmented/", "chunk") == "abc123"
    assert F.scene_name_of("/x/abc123.scene", "/d/augmented", "scene") == "abc123"
    with pytest.raises(NotImplementedError):
        F.scene_name_of("/x/a.scene", "/d/other", "scene")
