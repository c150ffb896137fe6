"""RPN proposal stage ("rpn3d").

`proposal_layer(...)` keeps the reference's 14-argument signature and return value
(lib/layer_utils/proposal_layer.py:11-15,204); `rpn_proposals(...)` is the sync-free form the forward
uses (logits in, padded device tensors + device count out).  Both run sis3d_rpn_proposals:
inside filter -> objectness -> exact stable top-N -> decode/clip -> NMS -> top-N, all on the GPU.
"""
import ctypes as C

import torch

from lib import _sis3d as S
from lib.utils.config import cfg


def rpn_proposals(levels, scene_dims, cfg_key="TEST", want_order=False, out=None):
    """levels: list of dict(cls, deltas, sizes [A,3] cuda f32, grid (gx,gy,gz), A, cls_mode).
    Returns rois [post,6], scores [post], level_ids int32 [post], num int32 [1] (+ order int32 [pre])."""
    pre, post = int(cfg[cfg_key].RPN_PRE_NMS_TOP_N), int(cfg[cfg_key].RPN_POST_NMS_TOP_N)
    thresh = float(cfg[cfg_key].RPN_NMS_THRESH)
    if pre <= 0 or post <= 0:
        raise S.Sis3dError("RPN_PRE/POST_NMS_TOP_N must be positive")
    dev = levels[0]["cls"].device
    arr = (S.RpnLevel * len(levels))()
    for i, lv in enumerate(levels):
        arr[i].cls, arr[i].deltas = lv["cls"].data_ptr(), lv["deltas"].data_ptr()
        arr[i].anchor_sizes = lv["sizes"].data_ptr()
        for k in range(3):
            arr[i].grid[k] = int(lv["grid"][k])
        arr[i].num_anchors, arr[i].cls_mode = int(lv["A"]), int(lv.get("cls_mode", 0))
        arr[i].cls_ld, arr[i].deltas_ld = int(lv.get("cls_ld", 0)), int(lv.get("deltas_ld", 0))
    nbytes = int(S.lib.sis3d_rpn_workspace_bytes(arr, len(levels), pre))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if out is not None:  # caller-provided outputs (views of one packed result buffer)
        rois, scores, lvl, num = out
    else:
        rois = torch.empty(post, 6, dtype=torch.float32, device=dev)
        scores = torch.empty(post, dtype=torch.float32, device=dev)
        lvl = torch.empty(post, dtype=torch.int32, device=dev)
        num = torch.empty(1, dtype=torch.int32, device=dev)
    order = torch.empty(pre, dtype=torch.int32, device=dev) if want_order else None
    S.check(S.lib.sis3d_rpn_proposals(arr, len(levels), 4, int(scene_dims[0]), int(scene_dims[1]), int(scene_dims[2]),
                                      int(cfg.ALLOW_BORDER), pre, post, S.f32(thresh), S.ptr(rois), S.ptr(scores),
                                      S.ptr(lvl), S.ptr(num), S.ptr(order), S.ptr(ws), C.c_size_t(nbytes), S.stream()),
            "rpn_proposals")
    return (rois, scores, lvl, num, order) if want_order else (rois, scores, lvl, num)


def proposal_layer(rpn_cls_prob_level1, rpn_bbox_pred_level1, all_anchors_level1,
                   rpn_cls_prob_level2, rpn_bbox_pred_level2, all_anchors_level2,
                   rpn_cls_prob_level3, rpn_bbox_pred_level3, all_anchors_level3,
                   scene_info, cfg_key, anchors_filter_level1, anchors_filter_level2, anchors_filter_level3):
    """Reference-compatible entry: probabilities [1,2,X,Y,Z,A], deltas [1,X,Y,Z,6A], anchors [K*A,6]."""
    if any(f is not None for f in (anchors_filter_level1, anchors_filter_level2, anchors_filter_level3)):
        raise NotImplementedError("anchor filters (FILTER_ANCHOR_LEVEL*) are a training-time feature")
    levels = []
    for prob, bbox, anchors, A in ((rpn_cls_prob_level1, rpn_bbox_pred_level1, all_anchors_level1, cfg.NUM_ANCHORS_LEVEL1),
                                   (rpn_cls_prob_level2, rpn_bbox_pred_level2, all_anchors_level2, cfg.NUM_ANCHORS_LEVEL2),
                                   (rpn_cls_prob_level3, rpn_bbox_pred_level3, all_anchors_level3, cfg.NUM_ANCHORS_LEVEL3)):
        if A == 0:
            continue
        if prob.shape[0] != 1:
            raise S.Sis3dError("proposal_layer: batch size 1 only (as the reference's RoI pooling)")
        a = anchors[:A].float()
        sizes = (a[:, 3:6] - a[:, 0:3]).contiguous().to(prob.device)
        levels.append(dict(cls=prob[0, 1].contiguous().float(), deltas=bbox[0].contiguous().float(), sizes=sizes,
                           grid=tuple(prob.shape[2:5]), A=A, cls_mode=1))
    rois, scores, lvl, num = rpn_proposals(levels, scene_info, cfg_key)
    n = int(num.item())
    return [rois[:n]], [scores[:n].view(-1, 1)], [lvl[:n].float()]
