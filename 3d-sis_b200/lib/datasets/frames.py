"""2-D inputs of a scene from the ScanNet frame folders (reference: lib/datasets/dataset.py:132-185,
211-262): depth PNGs (uint16 millimetres), 4x4 pose text files and colour frames under
`<BASE_IMAGE_PATH>/<scene>/{depth,pose,color}/<frameid>.*`, resized with nearest-neighbour sampling to a fixed
height and centre-cropped to the network's input size.

`FrameFolders` is a `view_provider` for lib.datasets.dataset.Dataset.  The resize/crop is pure index arithmetic
(no PIL object per frame): PIL's NEAREST rule (source coordinate accumulated in half-step-offset increments, truncated),
shifted by the centre-crop offset torchvision uses (round-half-even of the margin / 2);
tests/test_frames.py checks it against the reference's own torchvision/PIL calls."""
from __future__ import annotations

import math
import os

import numpy as np

from lib.utils.config import cfg


def load_pose(filename):
    """4 lines x 4 numbers -> float32 [4,4] (dataset.py:231-236)."""
    with open(filename) as f:
        lines = f.read().splitlines()
    if len(lines) != 4:
        raise ValueError(f"{filename}: a pose file has 4 lines")
    return np.asarray([ln.split()[:4] for ln in lines]).astype(np.float32)


def resize_crop_index(src_hw, new_dims):
    """Row / column source indices of `resize to height new_h (nearest) + centre crop to new_w` for an image of
    src_hw = (H, W); new_dims = [new_w, new_h] as in cfg.DEPTH_SHAPE / cfg.IMAGE_SHAPE (dataset.py:238-246)."""
    H, W = int(src_hw[0]), int(src_hw[1])
    new_w, new_h = int(new_dims[0]), int(new_dims[1])
    if [W, H] == [new_w, new_h]:
        return np.arange(H), np.arange(W)
    resize_w = int(math.floor(new_h * float(W) / float(H)))
    if resize_w < new_w:
        raise ValueError(f"image {W}x{H} is too narrow for a {new_w}x{new_h} centre crop (the reference would zero-pad)")
    rows = _pil_nearest_index(H, new_h)
    left = int(round((resize_w - new_w) / 2.0))  # torchvision center_crop; Python round = half to even
    cols = _pil_nearest_index(W, resize_w)[left:left + new_w]
    return rows, cols


def _pil_nearest_index(n_in, n_out):
    """Source index of every output pixel under PIL's NEAREST resize: the source coordinate starts at half a step and is
    ACCUMULATED step by step in double precision, then truncated (libImaging Geometry.c, ImagingScaleAffine -- the path
    8-bit colour frames and the int32 depth images scipy.misc.imread hands the reference take; PIL's native "I;16" mode
    would go through the generic transform, which multiplies instead of accumulating and differs in a few columns)."""
    step = float(n_in) / float(n_out)
    coord = np.cumsum(np.concatenate([[0.5 * step], np.full(n_out - 1, step)]))  # sequential sums, like the C loop
    return np.minimum(coord.astype(np.int64), n_in - 1)


def resize_crop_image(image, new_dims):
    rows, cols = resize_crop_index(image.shape[:2], new_dims)
    return image[rows][:, cols]


def _imread(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.array(im)


def load_depth(path, image_dims):
    """uint16 millimetres -> float32 metres at image_dims = [w, h] (dataset.py:248-253)."""
    return resize_crop_image(_imread(path), image_dims).astype(np.float32) / 1000.0


def load_image(path, image_dims, mean=None, std=None):
    """Colour frame -> float32 [3,h,w], /255 then (x - mean) / std per channel; label image -> [h,w] (dataset.py:255-266)."""
    image = resize_crop_image(_imread(path), image_dims)
    if image.ndim == 2:
        return image
    mean = np.asarray(cfg.COLOR_MEAN if mean is None else mean, dtype=np.float32).reshape(3, 1, 1)
    std = np.asarray(cfg.COLOR_STD if std is None else std, dtype=np.float32).reshape(3, 1, 1)
    x = np.transpose(image[..., :3], (2, 0, 1)).astype(np.float32) / np.float32(255.0)
    return (x - mean) / std


def scene_name_of(scene_path, base_image_path, mode):
    """Folder name of a .scene/.chunk file's frames (dataset.py:144-149)."""
    base = os.path.basename(scene_path)
    root = base_image_path.rstrip("/")
    if root.endswith("augmented"):
        return base.rsplit("_", 1)[0] if mode == "chunk" else base.split(".")[0]
    if root.endswith("square"):
        return base.split("__")[0]
    raise NotImplementedError("BASE_IMAGE_PATH must end in 'square' or 'augmented' (dataset.py:144-149)")


class FrameFolders:
    """view_provider(scene_path, frame_ids, world2grid, volume_dims) -> dict(images, depths, poses, world2grid, frameids).

    chunk mode: the frame ids and world2grid stored in the .chunk file are used; scene / benchmark mode: every depth
    frame of the folder and the folder's world2grid.txt minus the (10, 16, 10) volume padding (dataset.py:151-160).
    `features` (optional callable: float32 [n,3,H,W] -> [n,C,h,w], e.g. an ENet encoder) turns the colour frames into
    the features the 3-D network consumes; without it the normalised RGB frames are returned."""

    def __init__(self, base_image_path=None, mode="chunk", features=None):
        self.base = base_image_path if base_image_path is not None else cfg.BASE_IMAGE_PATH
        self.mode = mode
        self.features = features

    def __call__(self, scene_path, frame_ids, world2grid, volume_dims=None):
        d = os.path.join(self.base, scene_name_of(scene_path, self.base, self.mode))
        if self.mode != "chunk":
            frame_ids = sorted((f.split(".")[0] for f in os.listdir(os.path.join(d, "depth"))), key=lambda s: (len(s), s))
            world2grid = load_pose(os.path.join(d, "world2grid.txt"))
            world2grid[0:3, 3] -= np.array([10, 16, 10], dtype=np.float32)
        ext, kind = str(cfg.get("IMAGE_EXT", ".jpg")), str(cfg.get("IMAGE_TYPE", "color"))
        depths, images, poses, files = [], [], [], []
        for fid in frame_ids:
            poses.append(load_pose(os.path.join(d, "pose", f"{fid}.txt")))
            depths.append(load_depth(os.path.join(d, "depth", f"{fid}.png"), cfg.DEPTH_SHAPE))
            files.append(os.path.join(d, kind, f"{fid}{ext}"))
            images.append(load_image(files[-1], cfg.IMAGE_SHAPE))
        out = {"depths": np.stack(depths) if depths else np.zeros((0, cfg.DEPTH_SHAPE[1], cfg.DEPTH_SHAPE[0]), np.float32),
               "poses": np.stack(poses) if poses else np.zeros((0, 4, 4), np.float32),
               "world2grid": np.asarray(world2grid, dtype=np.float32), "frameids": list(frame_ids), "image_files": files}
        imgs = np.stack(images) if images else np.zeros((0, 3, cfg.IMAGE_SHAPE[1], cfg.IMAGE_SHAPE[0]), np.float32)
        out["images"] = self.features(imgs) if self.features is not None else imgs
        return out
