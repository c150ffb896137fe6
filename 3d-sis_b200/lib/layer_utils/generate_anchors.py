"""Anchor tables (reference: lib/layer_utils/generate_anchors.py:4-119).

The forward never materialises the K*A anchor list: csrc/rpn.cu rebuilds each anchor from its flat
index and the per-level size table.  `generate_anchors` is kept for API compatibility."""
import os

import numpy as np

from lib.utils.config import cfg


def read_anchor_sizes(name):
    """[A,3] float64 (w,h,l) rows of experiments/anchors/<name>."""
    path = name if os.path.isabs(name) else os.path.join(cfg.ANCHOR_DIR, name)
    with open(path) as f:
        rows = [[float(t) for t in line.split(",")] for line in f if line.strip()]
    return np.asarray(rows, dtype=np.float64).reshape(-1, 3)


def _tile(sizes, grid, stride):
    half = sizes / 2.0
    base = np.concatenate([-half, half], 1)                       # [A,6]
    idx = np.indices(tuple(int(g) for g in grid)).reshape(3, -1).T * float(stride)   # (x,y,z), z fastest
    return (idx[:, None, [0, 1, 2, 0, 1, 2]] + base[None]).reshape(-1, 6).astype(np.float32)


def generate_anchors(size_level1, size_level2, size_level3, feat_stride):
    out = []
    for lvl, size in enumerate((size_level1, size_level2, size_level3), 1):
        if cfg["NUM_ANCHORS_LEVEL%d" % lvl] != 0:
            out.append(_tile(read_anchor_sizes(cfg["ANCHORS_TYPE_LEVEL%d" % lvl]), size, feat_stride[lvl - 1]))
        else:
            out.append(None)
    return tuple(out)
