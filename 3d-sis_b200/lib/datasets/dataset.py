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


def load_mapping(label_file):
    """nyu40 label csv -> (raw id -> consecutive id, class weights with the background weight first) (dataset.py:268-283)."""
    import csv
    mapping, pre = {}, {}
    with open(label_file) as f:
        for row in csv.DictReader(f, delimiter=","):
            mapping[int(row["nyu40id"])] = int(row["mappedIdConsecutive"])
            pre[int(row["mappedIdConsecutive"])] = float(row["weight"])
    return mapping, [0.3280746813009404] + [pre[k] for k in sorted(pre)]


def part_in_chunk(b):
    """Fraction of a box inside the 96x48x96 chunk volume (dataset.py:219-229: the stored value is recomputed in chunk mode)."""
    overall = (b[3] - b[0]) * (b[4] - b[1]) * (b[5] - b[2])
    lo = [min(max(b[0], 0), 96), min(max(b[1], 0), 48), min(max(b[2], 0), 96)]
    hi = [min(max(b[3], 0), 96), min(max(b[4], 0), 48), min(max(b[5], 0), 96)]
    return (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]) / overall


class Dataset(torch.utils.data.Dataset):
    def __init__(self, data_location, mode="test", view_provider=None, label_mapping=None, label_weights=None):
        if isinstance(data_location, (list, tuple)):
            self.scenes = list(data_location)
        elif os.path.isdir(data_location):
            self.scenes = [os.path.join(data_location, x) for x in os.listdir(data_location)
                           if os.path.isfile(os.path.join(data_location, x))]
        else:
            with open(data_location) as f:
                self.scenes = [ln.strip() for ln in f if ln.strip()]
        self.mode = mode
        self.view_provider = view_provider
        self.mapping, self.weights = label_mapping, label_weights
        lm = str(cfg.get("LABEL_MAP", "") or "")
        if self.mapping is None and lm and os.path.isfile(lm):  # the reference always loads cfg.LABEL_MAP (dataset.py:39-40)
            self.mapping, self.weights = load_mapping(lm)

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
            gt_box[:, 6] = [self.mapping[int(v)] for v in gt_box[:, 6]]  # unknown raw labels raise, as in the reference
        masks = []
        for _, m in s["masks"]:
            m = m.astype(np.uint8)
            m[m > 1] = 0
            masks.append(m)
        if (cfg.get("KEEP_THRESH", 0.0) or cfg.USE_IMAGES) and s["part_in_volume"] is not None:
            # box filter (dataset.py:106-130): fraction inside the volume >= KEEP_THRESH and a non-zero class weight
            stats = [part_in_chunk(gt_box[i]) if self.mode == "chunk" else float(s["part_in_volume"][i])
                     for i in range(len(s["part_in_volume"]))]
            kept = [i for i in range(len(stats)) if stats[i] >= float(cfg.get("KEEP_THRESH", 0.0))
                    and (self.weights is None or self.weights[int(gt_box[i, 6])] != 0)]
            gt_box = gt_box[kept] if kept else np.zeros((0, 7), np.float32)
            if cfg.USE_MASK:
                masks = [masks[i] for i in kept]
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
