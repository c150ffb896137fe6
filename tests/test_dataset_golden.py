"""CPU: SURVEY row f3 pinned -- this repo's `.chunk` reader + frame loader against the item dict the UNMODIFIED reference
reader (lib/datasets/dataset.py:45-218, BinaryReader.py:10-36) produced from the very same bytes (tests/golden/dataset/,
generator oracle/make_golden_dataset.py): TSDF encoding, box floor/ceil + label mapping, the KEEP_THRESH / class-weight box
filter, mask clearing, max-height crop, world2grid inversion, frame ids, depth / colour resize-crop-normalise, poses."""
import os

import numpy as np

from conftest import GOLDEN


def test_chunk_reader_and_frame_loader_match_reference_reader():
    from lib.datasets.dataset import Dataset, collate_fn
    from lib.datasets.frames import FrameFolders
    from lib.utils.config import cfg, cfg_reset
    d = os.path.join(GOLDEN, "dataset")
    g = dict(np.load(os.path.join(d, "reference_item.npz")))
    cfg_reset()
    cfg.USE_IMAGES, cfg.USE_IMAGES_GT, cfg.USE_MASK = True, False, True
    cfg.KEEP_THRESH, cfg.TRUNCATED = float(g["keep_thresh"]), 3.0
    cfg.LABEL_MAP = os.path.join(d, "labels.csv")
    cfg.BASE_IMAGE_PATH = os.path.join(d, "frames_square")
    cfg.IMAGE_TYPE, cfg.IMAGE_EXT = "color", ".jpg"
    cfg.IMAGE_SHAPE, cfg.DEPTH_SHAPE = [328, 256], [41, 32]
    cfg.COLOR_MEAN, cfg.COLOR_STD = [0.496342, 0.466664, 0.440796], [0.277856, 0.28623, 0.291129]
    ds = Dataset([os.path.join(d, "sample__0.chunk")], "chunk", view_provider=FrameFolders(mode="chunk"))
    item = ds[0]
    assert np.array_equal(item["data"], g["data"]), "TSDF encoding / height crop"
    assert np.array_equal(np.asarray(item["gt_box"], dtype=np.float32), g["gt_box"]), "boxes: floor/ceil, label map, keep filter"
    assert len(item["gt_mask"]) == int(g["n_mask"])
    for j, m in enumerate(item["gt_mask"]):
        assert np.array_equal(m, g[f"mask_{j}"])
    v = item["nearest_images"]
    assert [int(i) for i in v["frameids"]] == list(g["frameids"])
    np.testing.assert_allclose(v["world2grid"], g["world2grid"], rtol=1e-6, atol=1e-6)
    assert np.array_equal(v["poses"], g["poses"])
    assert np.array_equal(v["depths"], g["depths"]), "depth: nearest resize + centre crop + mm -> m"
    np.testing.assert_allclose(v["images"], g["images"], atol=1e-6)  # /255, (x - mean) / std in fp32
    blobs = collate_fn([item])
    assert tuple(blobs["data"].shape) == (1,) + g["data"].shape and blobs["nearest_images"]["images"][0].shape == g["images"].shape
    cfg_reset()
