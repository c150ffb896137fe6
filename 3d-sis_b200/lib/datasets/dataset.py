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
    def __init__(self, data_location, mode="test", view_provider=None, label_mapping=None, label_weights=None,
                 device_decode=False):
        """device_decode=True: the voxel block is NOT transposed / TSDF-encoded on the host (four full-volume numpy passes,
        ~3 ms per chunk -- more than the whole GPU forward); the item carries the raw float32 block as stored in the file
        ('sdf_raw', pinned) and `collate_fn` turns it into the network input on the device with sis3d_chunk_decode."""
        self.device_decode = bool(device_decode)
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
        s = read_scene(path, raw_sdf=self.device_decode)
        data = None if self.device_decode else encode_tsdf(s["sdf"], float(cfg.TRUNCATED))
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
        if self.device_decode:
            X, Y, Z = s["dims"]
            vol_dims = (X, min(Y, max_h), Z)
            item = {"id": path, "sdf_raw": torch.from_numpy(np.array(s["sdf"], dtype=np.float32)), "sdf_dims": (X, Y, Z),
                    "y_keep": max_h, "gt_box": gt_box, "gt_mask": masks, "nearest_images": {}, "image_files": []}
            if torch.cuda.is_available():
                item["sdf_raw"] = item["sdf_raw"].pin_memory()
        else:
            data = data[:, :, :max_h, :]
            vol_dims = data.shape[1:]
            item = {"id": path, "data": np.ascontiguousarray(data), "gt_box": gt_box, "gt_mask": masks,
                    "nearest_images": {}, "image_files": []}
        if cfg.USE_IMAGES:
            if self.view_provider is None:
                raise RuntimeError("USE_IMAGES=True needs a view_provider (depth/pose/features source)")
            item["nearest_images"] = self.view_provider(path, s["frame_ids"], s["world2grid"], vol_dims)
        return item


def decode_on_device(sdf_raw, dims, y_keep, device=None):
    """Raw x-fastest float32 block of a .scene/.chunk file -> network input [1,2,X,Y',Z] on the device (sis3d_chunk_decode)."""
    import ctypes as C
    from lib import _sis3d as S
    X, Y, Z = (int(v) for v in dims)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    raw = sdf_raw.to(dev, non_blocking=True)
    Yk = min(Y, int(y_keep))
    data = torch.empty(1, 2, X, Yk, Z, dtype=torch.float32, device=dev)
    S.check(S.lib.sis3d_chunk_decode(S.ptr(raw), X, Y, Z, int(y_keep), C.c_float(float(cfg.TRUNCATED)), S.ptr(data), S.stream()),
            "chunk_decode")
    raw.record_stream(torch.cuda.current_stream())
    return data


def collate_fn(batch):
    """batch size 1 collate matching the blobs layout Network.forward expects (dataloader.py:8-44)."""
    assert len(batch) == 1, "the inference path processes one scene at a time"
    b = batch[0]
    if "sdf_raw" in b:  # device decode: raw block H2D + one kernel on the current stream (csrc/io.cu)
        data = decode_on_device(b["sdf_raw"], b["sdf_dims"], b["y_keep"])
    else:
        data = torch.from_numpy(b["data"]).unsqueeze(0)
    blobs = {"id": [b["id"]], "data": data, "gt_box": [torch.from_numpy(b["gt_box"])],
             "gt_mask": [b["gt_mask"]], "image_files": b["image_files"]}
    if b["nearest_images"]:
        v = b["nearest_images"]
        blobs["nearest_images"] = {"images": [torch.as_tensor(v["images"])], "depths": [torch.as_tensor(v["depths"])],
                                   "poses": [torch.as_tensor(v["poses"])], "world2grid": [torch.as_tensor(v["world2grid"])]}
    return blobs
