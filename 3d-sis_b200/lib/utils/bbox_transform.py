"""Box decode / clip utilities with the reference's signatures (lib/utils/bbox_transform.py:4-21,
59-99).  The hot path runs these inside the CUDA kernels (csrc/rpn.cu, csrc/roi.cu); the tensor
versions here serve callers that post-process on the host (driver code, tests)."""
import torch


def clip_boxes(boxes, scene_shape):
    lim = torch.tensor([float(scene_shape[i % 3]) for i in range(6)], dtype=boxes.dtype, device=boxes.device)
    return torch.minimum(boxes.clamp(min=0), lim)


def bbox_transform_inv(boxes, deltas):
    if len(boxes) == 0:
        return deltas.detach() * 0
    size = boxes[:, 3:6] - boxes[:, 0:3]
    ctr = boxes[:, 0:3] + 0.5 * size
    d = deltas.reshape(deltas.shape[0], -1, 6)
    pc = d[:, :, 0:3] * size.unsqueeze(1) + ctr.unsqueeze(1)
    ps = torch.exp(d[:, :, 3:6]) * size.unsqueeze(1)
    lo, hi = pc - 0.5 * ps, pc + 0.5 * ps
    # reference column order: all x-lo of every class, then y-lo ..., i.e. cat along dim 1 per coordinate
    return torch.cat([lo[:, :, 0], lo[:, :, 1], lo[:, :, 2], hi[:, :, 0], hi[:, :, 1], hi[:, :, 2]], 1)
