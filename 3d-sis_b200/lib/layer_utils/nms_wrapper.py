"""`nms(dets, thresh)` with the reference's signature (lib/layer_utils/nms_wrapper.py:7-16) on top
of libsis3d's on-device greedy NMS.  CUDA tensors only -- there is no CPU path."""
import ctypes as C

import torch

from lib import _sis3d as S


def nms(dets, thresh, return_count=False):
    """dets: cuda float32 [N,6] sorted by descending score -> LongTensor (cuda) of kept indices."""
    if not dets.is_cuda:
        raise S.Sis3dError("nms: expected a CUDA tensor (the B200 build has no CPU fallback)")
    dets = dets.contiguous().float()
    n = dets.shape[0]
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=dets.device)
    num = torch.zeros(1, dtype=torch.int32, device=dets.device)
    ws = torch.empty(max(int(S.lib.sis3d_nms_workspace_bytes(n)), 8), dtype=torch.uint8, device=dets.device)
    S.check(S.lib.sis3d_nms(S.ptr(dets), C.c_int(n), S.f32(thresh), S.ptr(keep), S.ptr(num), S.ptr(ws), S.stream()), "nms")
    if return_count:
        return keep, num
    return keep[:int(num.item())].contiguous()
