"""Scene/chunk dataset with the reference's item contract (lib/datasets/dataset.py:45-218): each item is
a dict with 'id', 'data' [2,X,Y,Z], 'gt_box' [n,7], 'gt_mask', 'nearest_images', 'image_files'.

The voxel part is parsed by lib/datasets/scene_io.py.  2-D inputs (depth maps, poses and either RGB
frames for ENet or ready ENet features) come from `view_provider(scene_path, frame_ids, world2grid)`,
because the image folders (BASE_IMAGE_PATH) are dataset-specific; a provider for synthetic data lives
in sis3d_synth."""
from __future__ import annotations

import os

import numpy as np
import torch

from lib.datasets.scene_io import encode_tsdf, read_scene
from lib.utils.config import cfg


class Dataset(torch.utils.data.Dataset):
    def __init__(self, data_location, mode="test", view_provider=None, label_mapping=None):
        if isinstance(data_location, (list, tuple)):
            self.scenes = list(data_location)
        else:
            with open(data_location) as f:
                self.scenes = [ln.strip() for ln in f if ln.strip()]
        self.mode = mode
        self.view_provider = view_provider
        self.mapping = label_mapping

    def __len__(self):
        return len(self.scenes)

    def __getitem__(self, idx):
        path = self.scenes[idx]
        s = read_scene(path)
        data = encode_tsdf(s["sdf"], float(cfg.TRUNCATED))
        gt_box = s["boxes"].copy()
        gt_box[:, 0:3] = np.floor(gt_box[:, 0:3])
        gt_box[:, 3:6] = np.ceil(gt_box[:, 3:6])
        if self.mapping is not None:
            gt_box[:, 6] = [self.mapping.get(int(v), 0) for v in gt_box[:, 6]]
        masks = []
        for _, m in s["masks"]:
            m = m.astype(np.uint8)
            m[m > 1] = 0
            masks.append(m)
        max_h = 480 if self.mode == "benchmark" else 48  # dataset.py:192-205
        keep = [i for i, b in enumerate(gt_box) if b[1] <= max_h and b[4] <= max_h]
        gt_box = gt_box[keep] if len(keep) else np.zeros((0, 7), np.float32)
        masks = [masks[i] for i in keep if i < len(masks)]
        data = data[:, :, :max_h, :]
        item = {"id": path, "data": np.ascontiguousarray(data), "gt_box": gt_box, "gt_mask": masks,
                "nearest_images": {}, "image_files": []}
        if cfg.USE_IMAGES:
            if self.view_provider is None:
                raise RuntimeError("USE_IMAGES=True needs a view_provider (depth/pose/features source)")
            item["nearest_images"] = self.view_provider(path, s["frame_ids"], s["world2grid"], data.shape[1:])
        return item


def collate_fn(batch):
    """batch size 1 collate matching the blobs layout Network.forward expects (dataloader.py:8-44)."""
    assert len(batch) == 1, "the inference path processes one scene at a time"
    b = batch[0]
    blobs = {"id": [b["id"]], "data": torch.from_numpy(b["data"]).unsqueeze(0), "gt_box": [torch.from_numpy(b["gt_box"])],
             "gt_mask": [b["gt_mask"]], "image_files": b["image_files"]}
    if b["nearest_images"]:
        v = b["nearest_images"]
        blobs["nearest_images"] = {"images": [torch.as_tensor(v["images"])], "depths": [torch.as_tensor(v["depths"])],
                                   "poses": [torch.as_tensor(v["poses"])], "world2grid": [torch.as_tensor(v["world2grid"])]}
    return blobs
