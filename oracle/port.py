"""oracle/port.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy + torch-CPU fp32) of the 3D-SIS dense-voxel TEST forward, the checker for
the sm_100a path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product path never does.

Pinning: the reference repository ships no golden vectors or tests, so this
port is pinned against outputs of the reference's own Python files executed in the build container
(oracle/make_golden.py -> tests/golden/*.npz; tests/test_oracle_golden.py).

Every function cites the reference lines it follows (paths relative to the reference root).
"""
from __future__ import annotations

import ctypes
import math
import os
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_ANCHOR_DIR = os.path.join(os.path.dirname(_HERE), "3d-sis_b200", "experiments", "anchors")


# ------------------------------------------------------------------------------------------------
# configuration (hot-path keys of lib/utils/config.py:16-247 + experiments/cfgs/*/rpn_class_mask_5.yml)
# ------------------------------------------------------------------------------------------------
def make_cfg(name="scannet", **over):
    c = SimpleNamespace(
        NET="ScanNet_Backbone", NUM_CLASSES=19, NUM_ANCHORS_LEVEL1=3, NUM_ANCHORS_LEVEL2=11,
        ANCHORS_TYPE_LEVEL1="scannet14_3.txt", ANCHORS_TYPE_LEVEL2="scannet14_11.txt",
        RPN_PRE_NMS_TOP_N=400, RPN_POST_NMS_TOP_N=200, RPN_NMS_THRESH=0.1, RPN_CHANNELS=256,
        CLASS_POOLING_SIZE=4, CLASS_THRESH=0.5, MASK_THRESH=0.5, USE_IMAGES=True, USE_MASK=True,
        USE_CLASS=True, VOXEL_SIZE=0.046875, PROJ_DEPTH_MIN=0.1, PROJ_DEPTH_MAX=4.0,
        DEPTH_SHAPE=[41, 32], NUM_IMAGE_CHANNELS=128, ALLOW_BORDER=0,
        INTRINSIC=[[37.01983, 0, 20, 0], [0, 38.52470, 15.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    if name == "suncg":
        c.NET, c.NUM_CLASSES = "SUNCG_Backbone", 26
        c.NUM_ANCHORS_LEVEL2 = 6
        c.ANCHORS_TYPE_LEVEL1, c.ANCHORS_TYPE_LEVEL2 = "suncg9_3.txt", "suncg9_6.txt"
        c.INTRINSIC = [[35.5070229, 0, 20, 0], [0, 36.9504013, 15.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    for k, v in over.items():
        setattr(c, k, v)
    return c


# ------------------------------------------------------------------------------------------------
# native helpers
# ------------------------------------------------------------------------------------------------
_lib = None


def _oracle_lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = ctypes.CDLL(path)
        _lib.oracle_nms3d.restype = ctypes.c_int
        _lib.oracle_roi_pool3d.restype = ctypes.c_int
    return _lib


def nms3d(boxes: np.ndarray, thresh: float, fma_mode: int = 0) -> np.ndarray:
    """Greedy 3D NMS on score-sorted boxes (lib/layer_utils/nms/pth_nms.py:7-45; CUDA variant
    lib/layer_utils/nms/src/cuda/nms_kernel.cu:11-79 + src/nms_cuda.c:35-62 when fma_mode=1)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    keep = np.zeros(max(n, 1), dtype=np.int64)
    k = _oracle_lib().oracle_nms3d(boxes.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n),
                                   ctypes.c_float(thresh), ctypes.c_int(fma_mode),
                                   keep.ctypes.data_as(ctypes.c_void_p))
    return keep[:k].copy()


def iou_matrix(boxes: np.ndarray, fma_mode: int = 0) -> np.ndarray:
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    out = np.zeros((n, n), dtype=np.float32)
    _oracle_lib().oracle_iou_matrix(boxes.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n),
                                    ctypes.c_int(fma_mode), out.ctypes.data_as(ctypes.c_void_p))
    return out


def roi_pool3d(feat: np.ndarray, rois: np.ndarray, pool=(4, 4, 4), scale=0.25):
    """3D RoI max pooling fwd (lib/layer_utils/roi_pooling/src/cuda/roi_pooling_kernel.cu:15-109,
    src/roi_pooling.c:6-124).  feat [C,W,H,L] or [1,C,W,H,L]; returns (out[n,C,p,p,p], argmax int32)."""
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    if feat.ndim == 5:
        feat = feat[0]
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    C, W, H, L = feat.shape
    n = rois.shape[0]
    out = np.zeros((n, C) + tuple(pool), dtype=np.float32)
    arg = np.zeros((n, C) + tuple(pool), dtype=np.int32)
    if n:
        _oracle_lib().oracle_roi_pool3d(feat.ctypes.data_as(ctypes.c_void_p), C, W, H, L,
                                        rois.ctypes.data_as(ctypes.c_void_p), n, pool[0], pool[1], pool[2],
                                        ctypes.c_float(scale), out.ctypes.data_as(ctypes.c_void_p),
                                        arg.ctypes.data_as(ctypes.c_void_p))
    return out, arg


# ------------------------------------------------------------------------------------------------
# anchors  (lib/layer_utils/generate_anchors.py:4-119)
# ------------------------------------------------------------------------------------------------
def read_anchor_table(fname):
    rows = []
    with open(os.path.join(_ANCHOR_DIR, fname)) as f:
        for line in f:
            if line.strip():
                rows.append([float(t) for t in line.strip().split(",")])
    return np.asarray(rows, dtype=np.float64)


def generate_anchors(level_size, table, stride=4):
    """[K*A,6] float32, order (x,y,z,a) with a fastest; centred on stride*index, no half-stride
    offset (generate_anchors.py:78-88)."""
    base = np.concatenate([-table / 2.0, table / 2.0], 1)  # [A,6] float64
    sx, sy, sz = (np.arange(0, s) * stride for s in level_size)
    gx, gy, gz = np.meshgrid(sx, sy, sz, indexing="ij")
    shifts = np.stack([gx.ravel(), gy.ravel(), gz.ravel()] * 2, 1).astype(np.float64)
    anchors = base[None, :, :] + shifts[:, None, :]
    return anchors.reshape(-1, 6).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# back-projection index build  (lib/layer_utils/projection.py:27-121)
# ------------------------------------------------------------------------------------------------
def _depth_to_skeleton(intr, ux, uy, depth):  # projection.py:16-19
    x = (ux - intr[0][2]) / intr[0][0]
    y = (uy - intr[1][2]) / intr[1][1]
    return torch.tensor([depth * x, depth * y, depth], dtype=torch.float32)


def frustum_bounds(cfg, world_to_grid, camera_to_world):
    """projection.py:27-49 (float32 torch ops in the same order)."""
    intr, (w, h) = cfg.INTRINSIC, cfg.DEPTH_SHAPE
    cp = torch.ones(8, 4, 1, dtype=torch.float32)
    for k, (ux, uy, d) in enumerate([(0, 0, cfg.PROJ_DEPTH_MIN), (w - 1, 0, cfg.PROJ_DEPTH_MIN),
                                     (w - 1, h - 1, cfg.PROJ_DEPTH_MIN), (0, h - 1, cfg.PROJ_DEPTH_MIN),
                                     (0, 0, cfg.PROJ_DEPTH_MAX), (w - 1, 0, cfg.PROJ_DEPTH_MAX),
                                     (w - 1, h - 1, cfg.PROJ_DEPTH_MAX), (0, h - 1, cfg.PROJ_DEPTH_MAX)]):
        cp[k, :3, 0] = _depth_to_skeleton(intr, ux, uy, d)
    p = torch.bmm(camera_to_world.repeat(8, 1, 1), cp)
    pl = torch.round(torch.bmm(world_to_grid.repeat(8, 1, 1), torch.floor(p)))
    pu = torch.round(torch.bmm(world_to_grid.repeat(8, 1, 1), torch.ceil(p)))
    bmin = torch.minimum(pl[:, :3, 0].min(0)[0], pu[:, :3, 0].min(0)[0])
    bmax = torch.maximum(pl[:, :3, 0].max(0)[0], pu[:, :3, 0].max(0)[0])
    return bmin, bmax


def compute_projection(cfg, depth, camera_to_world, world_to_grid, volume_dims, want_margin=False):
    """Voxel->pixel index lists for one view (projection.py:52-121).

    Returns None (view is "killed") or (lin3d int64[k], lin2d int64[k]) in ascending lin3d order,
    lin3d = z*X*Y + y*X + x, lin2d = v*W + u.  Integer floor division restates the torch-0.4
    LongTensor `/` the reference relies on (projection.py:68,70,80,82).
    With want_margin a third array gives, per *frustum* voxel, how far (in float units) its
    accept/reject decision is from flipping -- tests use it to exclude knife-edge voxels.
    """
    depth = torch.as_tensor(depth, dtype=torch.float32)
    c2w = torch.as_tensor(camera_to_world, dtype=torch.float32)
    w2g = torch.as_tensor(world_to_grid, dtype=torch.float32)
    X, Y, Z = (int(v) for v in volume_dims)
    W, H = cfg.DEPTH_SHAPE
    fx, fy, cx, cy = cfg.INTRINSIC[0][0], cfg.INTRINSIC[1][1], cfg.INTRINSIC[0][2], cfg.INTRINSIC[1][2]
    world_to_camera = torch.inverse(c2w)
    grid_to_world = torch.inverse(w2g)
    bmin, bmax = frustum_bounds(cfg, w2g, c2w)
    bmin = torch.clamp(bmin, min=0).float()
    bmax = torch.minimum(bmax, torch.tensor([X, Y, Z], dtype=torch.float32)).float()

    lin = torch.arange(0, X * Y * Z, dtype=torch.int64)
    cz = lin // (X * Y)
    tmp = lin - cz * (X * Y)
    cy_ = tmp // X
    cx_ = tmp % X
    coords = torch.stack([cx_.float(), cy_.float(), cz.float(), torch.ones(lin.numel())], 0)
    m = (coords[0] >= bmin[0]) & (coords[1] >= bmin[1]) & (coords[2] >= bmin[2]) & \
        (coords[0] < bmax[0]) & (coords[1] < bmax[1]) & (coords[2] < bmax[2])
    if not m.any():
        return None
    lin = lin[m]
    coords = coords[:, m]
    p = torch.mm(world_to_camera, torch.mm(grid_to_world, coords))
    p[0] = (p[0] * fx) / p[2] + cx
    p[1] = (p[1] * fy) / p[2] + cy
    pi = torch.round(p).long()
    valid = (pi[0] >= 0) & (pi[1] >= 0) & (pi[0] < W) & (pi[1] < H)
    if not valid.any():
        return None
    vx, vy = pi[0][valid], pi[1][valid]
    lin2d = vy * W + vx
    dvals = depth.reshape(-1)[lin2d]
    pz = p[2][valid]
    dm = (dvals >= cfg.PROJ_DEPTH_MIN) & (dvals <= cfg.PROJ_DEPTH_MAX) & ((dvals - pz).abs() <= cfg.VOXEL_SIZE)
    if not dm.any():
        return None
    out3, out2 = lin[valid][dm].numpy(), lin2d[dm].numpy()
    if not want_margin:
        return out3, out2
    # decision margins (only meaningful for tests)
    fr = lambda t: (t - torch.floor(t) - 0.5).abs()
    margin = torch.minimum(fr(p[0]), fr(p[1]))
    mz = torch.full_like(margin, 1e9)
    dv_all = torch.zeros_like(margin)
    dv_all[valid] = dvals
    mz[valid] = torch.minimum(((dvals - pz).abs() - cfg.VOXEL_SIZE).abs(),
                              torch.minimum((dvals - cfg.PROJ_DEPTH_MIN).abs(), (dvals - cfg.PROJ_DEPTH_MAX).abs()))
    margin = torch.minimum(margin, mz)
    return out3, out2, (lin.numpy(), margin.numpy())


def projection_scatter(feat, lin3d, lin2d, volume_dims):
    """Projection.forward (projection.py:129-136): [C,h,w] -> [C,Z,Y,X], zeros where uncovered."""
    feat = torch.as_tensor(feat, dtype=torch.float32)
    C = feat.shape[0]
    X, Y, Z = (int(v) for v in volume_dims)
    out = torch.zeros(C, Z * Y * X)
    if len(lin3d):
        out.index_copy_(1, torch.as_tensor(lin3d), feat.reshape(C, -1)[:, torch.as_tensor(lin2d)])
    return out.view(C, Z, Y, X)


def backproject_views(cfg, feats, depths, poses, world2grid, volume_dims):
    """trainval.py:797-820 + network.py:194-239: per-view index build, drop views with no valid
    projection (killing_inds), scatter + running max, permute to [1,C,X,Y,Z].

    NOTE (reference behaviour kept on purpose): the surviving index lists are stacked densely
    (trainval.py:805-820) while network.py:220-223 zips them against ALL feature maps and skips by
    position, so after a killed view k the later maps are paired with the *next* view's indices.
    """
    n = len(feats)
    maps = [compute_projection(cfg, depths[i], poses[i], world2grid, volume_dims) for i in range(n)]
    killing = [i for i, m in enumerate(maps) if m is None]
    real = [m for m in maps if m is not None]
    if not real:
        raise ValueError("no view has a valid projection (reference would fail in zip(*[]))")
    vol = None
    for counter, (ft, (l3, l2)) in enumerate(zip(feats, real)):
        if counter in killing:
            continue
        cur = projection_scatter(ft, l3, l2, volume_dims)
        vol = cur if vol is None else torch.maximum(vol, cur)
    if vol is None:
        raise ValueError("every paired view was skipped")
    return vol.permute(0, 3, 2, 1).unsqueeze(0).contiguous(), killing  # [1,C,X,Y,Z]


# ------------------------------------------------------------------------------------------------
# 3D backbone  (lib/nets/backbones.py:17-40, 98-113, 118-231)
# ------------------------------------------------------------------------------------------------
def _t(w, k):
    return torch.as_tensor(w[k])


def bottleneck(x, w, p):
    """backbones.py:28-40: 1x1 -> relu -> 3x3x3 -> relu -> 1x1, += residual, relu."""
    y = F.relu(F.conv3d(x, _t(w, p + ".conv1.weight"), _t(w, p + ".conv1.bias")))
    y = F.relu(F.conv3d(y, _t(w, p + ".conv2.weight"), _t(w, p + ".conv2.bias"), padding=1))
    y = F.conv3d(y, _t(w, p + ".conv3.weight"), _t(w, p + ".conv3.bias"))
    return F.relu(y + x)


def _seq(x, w, prefix, spec):
    for kind, idx in spec:
        name = f"{prefix}.{idx}"
        if kind == "k2s2":
            x = F.relu(F.conv3d(x, _t(w, name + ".weight"), None, stride=2))
        elif kind == "k3":
            x = F.relu(F.conv3d(x, _t(w, name + ".weight"), None, padding=1))
        elif kind == "bneck":
            x = bottleneck(x, w, name)
        elif kind == "pool":
            x = F.max_pool3d(x, 3, 1, 1)
    return x


_SPECS = {
    "ScanNet_Backbone": dict(
        geometry1=[("k2s2", 0), ("bneck", 2), ("bneck", 3), ("k2s2", 4), ("bneck", 6), ("bneck", 7)],
        color=[("k2s2", 0), ("bneck", 2), ("pool", 3), ("k2s2", 4), ("bneck", 6), ("pool", 7)],
        geometry2=[("k3", 0), ("bneck", 2), ("bneck", 3), ("pool", 4)]),
    "SUNCG_Backbone": dict(
        geometry1=[("k2s2", 0), ("bneck", 2), ("k2s2", 3), ("bneck", 5)],
        color=[("k2s2", 0), ("bneck", 2), ("k2s2", 3), ("bneck", 5)],
        geometry2=[("k3", 0), ("bneck", 2)]),
}


def backbone(cfg, w, scene, imageft):
    """Base_Backbone._backbone (backbones.py:98-113), image+geometry or geometry-only branch."""
    spec = _SPECS[cfg.NET]
    geo = _seq(scene, w, "geometry1", spec["geometry1"])
    if cfg.USE_IMAGES:
        col = _seq(imageft, w, "color", spec["color"])
        level1 = torch.cat([col, geo], 1)
    else:
        level1 = geo
    level2 = _seq(level1, w, "geometry2", spec["geometry2"])
    return level1, level2


# ------------------------------------------------------------------------------------------------
# RPN  (lib/nets/network.py:537-587, lib/layer_utils/proposal_layer.py:11-204,
#       lib/utils/bbox_transform.py:4-21,59-99)
# ------------------------------------------------------------------------------------------------
def rpn_heads(w, feat, lvl, A):
    """network.py:538-550: returns (fg prob [X,Y,Z,A] flattened, deltas [X*Y*Z*A,6])."""
    h = F.relu(F.conv3d(feat, _t(w, f"rpn_net_level{lvl}.weight"), _t(w, f"rpn_net_level{lvl}.bias"), padding=1))
    bbox = F.conv3d(h, _t(w, f"rpn_bbox_pred_net_level{lvl}.weight"), _t(w, f"rpn_bbox_pred_net_level{lvl}.bias"))
    bbox = bbox.permute(0, 2, 3, 4, 1).contiguous()
    cls = F.conv3d(h, _t(w, f"rpn_cls_score_net_level{lvl}.0.weight"), _t(w, f"rpn_cls_score_net_level{lvl}.0.bias"))
    B, _, X, Y, Z = cls.shape
    cls = cls.view(B, 2, A, X, Y, Z).permute(0, 1, 3, 4, 5, 2).contiguous()
    prob = F.softmax(cls, dim=1)  # implicit dim for 6-D input in torch 0.4 is 1
    return prob[0, 1].reshape(-1), bbox[0].reshape(-1, 6), h


def bbox_transform_inv(boxes, deltas):
    """bbox_transform.py:59-99 (float32, one rounding per op -- no FMA contraction on CPU)."""
    boxes, deltas = torch.as_tensor(boxes), torch.as_tensor(deltas)
    if len(boxes) == 0:
        return deltas * 0
    wd, ht, ln = boxes[:, 3] - boxes[:, 0], boxes[:, 4] - boxes[:, 1], boxes[:, 5] - boxes[:, 2]
    cx, cy, cz = boxes[:, 0] + 0.5 * wd, boxes[:, 1] + 0.5 * ht, boxes[:, 2] + 0.5 * ln
    px, py, pz = deltas[:, 0] * wd + cx, deltas[:, 1] * ht + cy, deltas[:, 2] * ln + cz
    pw, ph, pl = torch.exp(deltas[:, 3]) * wd, torch.exp(deltas[:, 4]) * ht, torch.exp(deltas[:, 5]) * ln
    return torch.stack([px - 0.5 * pw, py - 0.5 * ph, pz - 0.5 * pl,
                        px + 0.5 * pw, py + 0.5 * ph, pz + 0.5 * pl], 1)


def clip_boxes(boxes, dims):
    """bbox_transform.py:4-21: clamp to [0, dim] inclusive."""
    d = [float(v) for v in dims]
    return torch.stack([boxes[:, i].clamp(0, d[i % 3]) for i in range(6)], 1)


def inside_mask(anchors, dims, border=0):
    """proposal_layer.py:36-43: lo >= -border and hi < dim + border (strict)."""
    a = np.asarray(anchors)
    return ((a[:, 0] >= -border) & (a[:, 1] >= -border) & (a[:, 2] >= -border) &
            (a[:, 3] < dims[0] + border) & (a[:, 4] < dims[1] + border) & (a[:, 5] < dims[2] + border))


def proposal_layer(cfg, levels, dims, fma_mode=0):
    """proposal_layer.py:11-204 for batch 1.  levels = [(prob_flat, deltas[K*A,6], anchors[K*A,6]), ...]
    in level order.  Sort is descending and STABLE (ties -> lower concatenated index first); the
    reference's torch.sort leaves tie order unspecified.
    Returns dict(rois, scores, level_inds, order (pre-NMS top-N indices into the concatenated inside
    list), keep)."""
    props, scores, lvls = [], [], []
    for li, (prob, deltas, anchors) in enumerate(levels):
        inside = np.nonzero(inside_mask(anchors, dims, cfg.ALLOW_BORDER))[0]
        a = torch.as_tensor(anchors)[inside]
        d = torch.as_tensor(deltas)[inside]
        p = clip_boxes(bbox_transform_inv(a, d), dims) if len(inside) else torch.zeros(0, 6)
        props.append(p)
        scores.append(torch.as_tensor(prob)[inside])
        lvls.append(torch.full((len(inside),), float(li + 1)))
    props, scores, lvls = torch.cat(props, 0), torch.cat(scores, 0), torch.cat(lvls, 0)
    order = torch.sort(scores, descending=True, stable=True)[1]
    if cfg.RPN_PRE_NMS_TOP_N > 0:
        order = order[:cfg.RPN_PRE_NMS_TOP_N]
    p_sorted = props[order]
    keep = nms3d(p_sorted.numpy(), cfg.RPN_NMS_THRESH, fma_mode)
    if cfg.RPN_POST_NMS_TOP_N > 0:
        keep = keep[:cfg.RPN_POST_NMS_TOP_N]
    keep_t = torch.as_tensor(keep)
    return dict(rois=p_sorted[keep_t], scores=scores[order][keep_t], level_inds=lvls[order][keep_t],
                order=order.numpy(), keep=keep, sorted_boxes=p_sorted.numpy(), all_scores=scores.numpy())


# ------------------------------------------------------------------------------------------------
# RoI pooling + classifier  (network.py:503-534, 589-604; backbones.py:92-96)
# ------------------------------------------------------------------------------------------------
def roi_pool_levels(cfg, level_feats, rois, level_inds):
    P = cfg.CLASS_POOLING_SIZE
    n = rois.shape[0]
    C = level_feats[0].shape[1]
    pool5 = np.zeros((n, C, P, P, P), dtype=np.float32)
    for li, feat in enumerate(level_feats):
        sel = np.nonzero(np.asarray(level_inds) == li + 1)[0]
        if len(sel):
            out, _ = roi_pool3d(feat.numpy(), np.asarray(rois)[sel], (P, P, P), 0.25)
            pool5[sel] = out
    return pool5


def classifier(w, pool5):
    x = torch.as_tensor(pool5).reshape(pool5.shape[0], -1)
    for i in (0, 2, 4):
        x = F.relu(F.linear(x, _t(w, f"classifier.{i}.weight"), _t(w, f"classifier.{i}.bias")))
    cls_score = F.linear(x, _t(w, "classifier_cls_score_net.weight"), _t(w, "classifier_cls_score_net.bias"))
    bbox_pred = F.linear(x, _t(w, "classifier_bbox_pred_net.weight"), _t(w, "classifier_bbox_pred_net.bias"))
    return cls_score, cls_score.argmax(1), F.softmax(cls_score, dim=1), bbox_pred


# ------------------------------------------------------------------------------------------------
# mask branch  (network.py:283-317; backbones.py:236-287)
# ------------------------------------------------------------------------------------------------
def mask_head(w, crop):
    x = crop
    for i in (0, 2, 4, 6, 8):
        x = F.relu(F.conv3d(x, _t(w, f"mask_backbone.geometry.{i}.weight"), None, padding=1))
    x = F.conv3d(x, _t(w, "mask_backbone.geometry.10.weight"), None)
    return torch.sigmoid(x)


def detection_boxes(cfg, rois, cls_pred, cls_prob, bbox_pred, dims):
    """network.py:285-301 == trainval.py:825-858: class-specific decode, clip, conf filter,
    Python-3 round() (half-to-even) degenerate-box filter.  Returns (pred_box[n,6] f32, conf[n] f64,
    keep bool[n], crops int[n,6])."""
    n = rois.shape[0]
    cls_pred = np.asarray(cls_pred)
    box_reg = np.zeros((n, 6))
    conf = np.zeros(n)
    bp, cp = np.asarray(bbox_pred), np.asarray(cls_prob)
    for i in range(n):
        box_reg[i] = bp[i, cls_pred[i] * 6:(cls_pred[i] + 1) * 6]
        conf[i] = cp[i, cls_pred[i]]
    pred_box = clip_boxes(bbox_transform_inv(torch.as_tensor(rois), torch.from_numpy(box_reg).float()), dims).numpy() \
        if n else np.zeros((0, 6), np.float32)
    keep = conf > cfg.CLASS_THRESH
    crops = np.zeros((n, 6), dtype=np.int64)
    for i, b in enumerate(pred_box):
        r = [int(round(float(v))) for v in b]
        crops[i] = r
        if r[0] >= r[3] or r[1] >= r[4] or r[2] >= r[5]:
            keep[i] = False
    return pred_box, conf, keep, crops


# ------------------------------------------------------------------------------------------------
# the whole TEST forward  (lib/nets/network.py:187-317)
# ------------------------------------------------------------------------------------------------
def forward(cfg, w, data, views=None, timings=None, fma_mode=0):
    """data [1,2,X,Y,Z] float32; views = dict(feats, depths, poses, world2grid) when cfg.USE_IMAGES.
    Returns a dict mirroring Network._predictions (+ intermediates used by the parity tests)."""
    import time
    t0 = time.perf_counter()
    out = {}
    with torch.no_grad():
        scene = torch.as_tensor(data, dtype=torch.float32)
        dims = tuple(int(v) for v in scene.shape[2:])
        imageft = None
        if cfg.USE_IMAGES:
            imageft, killing = backproject_views(cfg, views["feats"], views["depths"], views["poses"],
                                                 views["world2grid"], dims)
            out["imageft"], out["killing_inds"] = imageft, killing
        t1 = time.perf_counter()
        level1, level2 = backbone(cfg, w, scene, imageft)
        out["level1"], out["level2"] = level1, level2
        t2 = time.perf_counter()
        levels = []
        for lvl, (feat, A, tab) in enumerate(((level1, cfg.NUM_ANCHORS_LEVEL1, cfg.ANCHORS_TYPE_LEVEL1),
                                              (level2, cfg.NUM_ANCHORS_LEVEL2, cfg.ANCHORS_TYPE_LEVEL2)), 1):
            prob, deltas, _ = rpn_heads(w, feat, lvl, A)
            anchors = generate_anchors(feat.shape[2:], read_anchor_table(tab), 4)
            levels.append((prob, deltas, anchors))
            out[f"rpn_prob_level{lvl}"], out[f"rpn_deltas_level{lvl}"] = prob, deltas
        prop = proposal_layer(cfg, levels, dims, fma_mode)
        out.update(rois=prop["rois"], roi_scores=prop["scores"], level_inds=prop["level_inds"],
                   rpn_order=prop["order"], nms_keep=prop["keep"], rpn_sorted_boxes=prop["sorted_boxes"],
                   rpn_all_scores=prop["all_scores"])
        t3 = time.perf_counter()
        if cfg.USE_CLASS:
            pool5 = roi_pool_levels(cfg, (level1, level2), prop["rois"].numpy(), prop["level_inds"].numpy())
            cls_score, cls_pred, cls_prob, bbox_pred = classifier(w, pool5)
            out.update(pool5=pool5, cls_score=cls_score, cls_pred=cls_pred, cls_prob=cls_prob, bbox_pred=bbox_pred)
        t4 = time.perf_counter()
        if cfg.USE_MASK and cfg.USE_CLASS:
            pred_box, conf, keep, crops = detection_boxes(cfg, prop["rois"].numpy(), cls_pred.numpy(),
                                                          cls_prob.numpy(), bbox_pred.numpy(), dims)
            masks = []
            for i in np.nonzero(keep)[0]:
                c = crops[i]
                masks.append(mask_head(w, scene[:, :, c[0]:c[3], c[1]:c[4], c[2]:c[5]]))
            out.update(pred_box=pred_box, pred_conf=conf, mask_keep=keep, mask_crops=crops, mask_pred=masks)
        t5 = time.perf_counter()
    if timings is not None:
        timings.update(project=t1 - t0, backbone=t2 - t1, rpn=t3 - t2, roi_cls=t4 - t3, mask=t5 - t4, total=t5 - t0)
    return out


# ------------------------------------------------------------------------------------------------
# Row f2 (next): 2-D ENet encoder, images [n,3,256,328] -> features [n,128,32,41]
# (lib/nets/enet.py:130-590 = create_enet modules 0..25; create_enet_for_3d, enet.py:697-715, keeps exactly those as
# model_fixed + model_trainable; network.py:199-213 feeds their output to the back-projection).
# Restated as a layer program; `params` = the tensors of the reference state_dict in its own order with the
# num_batches_tracked counters dropped (tests/golden/enet_encoder.npz carries them as p000, p001, ...).
# ------------------------------------------------------------------------------------------------
_ENET_STAGE23 = [("reg", 1), ("reg", 2), ("asym", 5), ("reg", 4), ("reg", 1), ("reg", 8), ("asym", 5), ("reg", 16)]
ENET_PROGRAM = ([("down", 16, 64, 0.01)] + [("reg", 16, 64, 0.01, 1)] * 4 + [("down", 32, 128, 0.1)] +
                [(kind, 32, 128, 0.1, arg) for kind, arg in _ENET_STAGE23 * 2])
ENET_BN_EPS = 1e-3  # nn.BatchNorm2d(C, 0.001, 0.1, True), enet.py:138


def enet_encoder(params, images):
    """Inference forward of the ENet encoder on CPU (eval-mode BatchNorm, Torch7-style Dropout2d that scales by
    (1 - p) at inference, enet.py:89-95).  params: list of float32 tensors in reference order."""
    it = iter(params)

    def bn_prelu(x, prelu=True):
        w, b, mean, var = next(it), next(it), next(it), next(it)
        x = F.batch_norm(x, mean, var, w, b, False, 0.1, ENET_BN_EPS)
        return F.prelu(x, next(it)) if prelu else x

    x = images
    w0, b0 = next(it), next(it)  # initial block: conv 3->13 k3 s2 p1 || 2x2 max-pool, concatenated (enet.py:132-137)
    x = torch.cat((F.conv2d(x, w0, b0, stride=2, padding=1), F.max_pool2d(x, 2, 2)), 1)
    x = bn_prelu(x)
    for op in ENET_PROGRAM:
        kind, mid, cout, p = op[0], op[1], op[2], op[3]
        skip = x
        if kind == "down":  # 2x2/s2 projection on the main branch, max-pool + zero channel padding on the skip
            y = F.conv2d(x, next(it), None, stride=2)
            skip = F.max_pool2d(x, 2, 2)
            skip = torch.cat((skip, skip.new_zeros(skip.shape[0], cout - skip.shape[1], *skip.shape[2:])), 1)
        else:
            y = F.conv2d(x, next(it), None)
        y = bn_prelu(y)
        if kind == "asym":  # 1x5 (no bias) then 5x1 (bias), enet.py asymmetric bottlenecks
            y = F.conv2d(y, next(it), None, padding=(0, 2))
            w, b = next(it), next(it)
            y = F.conv2d(y, w, b, padding=(2, 0))
        else:
            d = op[4] if kind == "reg" else 1
            w, b = next(it), next(it)
            y = F.conv2d(y, w, b, padding=d, dilation=d)
        y = bn_prelu(y)
        y = bn_prelu(F.conv2d(y, next(it), None), prelu=False) * (1.0 - p)
        x = F.prelu(y + skip, next(it))
    if next(it, None) is not None:
        raise ValueError("enet_encoder: unused parameters (not the encoder's state_dict?)")
    return x


# ------------------------------------------------------------------------------------------------
# Row f4 (next): voxel predictions -> mesh-vertex instance labels, literal restatement of the loops of
# tools/scannet_benchmark/vox2mesh.py:55-106 (small cases only: it is voxel-by-voxel / vertex-by-vertex Python)
# ------------------------------------------------------------------------------------------------
def vox2mesh_paint(pred_box, pred_class, pred_conf, pred_mask, dims):
    scene = np.zeros(dims)
    for box_ind, box in enumerate(pred_box):  # vox2mesh.py:55-69
        lo = [int(round(box[a])) for a in range(3)]
        hi = [int(round(box[a + 3])) for a in range(3)]
        for i in range(lo[0], hi[0]):
            for j in range(lo[1], hi[1]):
                for k in range(lo[2], hi[2]):
                    if pred_mask[box_ind][i - lo[0], j - lo[1], k - lo[2]] != 0 and scene[i, j, k] == 0:
                        scene[i, j, k] = box_ind * 100 + pred_class[box_ind] + pred_conf[box_ind] - 0.01
    return scene


def vox2mesh_labels(mesh_vertices, world2grid, scene):
    """-> (instance_class, instance_mask, instance_conf) dicts as vox2mesh.py:83-106 builds them; vertices whose
    3x3x3 neighbourhood leaves the volume are skipped (the reference would index out of range there)."""
    def nn_search(x, y, z):  # vox2mesh.py:71-81
        if scene[x, y, z] != 0:
            return x, y, z
        for i in [-1, 0, 1]:
            for j in [-1, 0, 1]:
                for k in [-1, 0, 1]:
                    if scene[x + i, y + j, z + k] != 0:
                        return x + i, y + j, z + k
        return -1, -1, -1

    instance_mask, instance_conf, instance_class = {}, {}, {}
    for ind, vertex in enumerate(mesh_vertices):
        g = np.round(np.matmul(world2grid, np.append(vertex, 1)))
        x, y, z = int(round(g[0])), int(round(g[1])), int(round(g[2]))
        if not (1 <= x <= scene.shape[0] - 2 and 1 <= y <= scene.shape[1] - 2 and 1 <= z <= scene.shape[2] - 2):
            continue
        x, y, z = nn_search(x, y, z)
        if x == -1:
            continue
        conf = np.modf(scene[x, y, z])[0]
        instance_id = int(int(scene[x, y, z]) / 100)
        class_id = int(scene[x, y, z]) % 100
        if instance_id not in instance_class:
            instance_class[instance_id] = class_id
            instance_mask[instance_id] = [ind]
            instance_conf[instance_id] = conf
        else:
            instance_mask[instance_id].append(ind)
    return instance_class, instance_mask, instance_conf
