"""`RoIPoolFunction(pw, ph, pl, scale)(features, rois)` with the reference's call form
(lib/layer_utils/roi_pooling/roi_pool.py:9-38) on libsis3d's forward kernel.  Inference only."""
import ctypes as C

import torch

from lib import _sis3d as S


class RoIPoolFunction:
    def __init__(self, pooled_width, pooled_height, pooled_length, spatial_scale):
        self.pooled_width, self.pooled_height, self.pooled_length = int(pooled_width), int(pooled_height), int(pooled_length)
        self.spatial_scale = float(spatial_scale)
        self.argmax = self.rois = self.feature_size = None

    def __call__(self, features, rois):
        return self.forward(features, rois)

    def forward(self, features, rois):
        """features [1,C,W,H,L] (reference NCDHW layout), rois [n,6] -> [n,C,pw,ph,pl]."""
        if features.dim() != 5 or features.shape[0] != 1:
            raise S.Sis3dError("RoIPoolFunction: features must be [1,C,W,H,L]")  # reference returns 0 silently
        if rois.dim() != 2 or rois.shape[1] != 6:
            raise S.Sis3dError("RoIPoolFunction: rois must be [n,6]")
        features, rois = features.contiguous().float(), rois.contiguous().float()
        _, Cn, W, H, L = features.shape
        n = rois.shape[0]
        out = torch.zeros(n, Cn, self.pooled_width, self.pooled_height, self.pooled_length, device=features.device)
        arg = torch.zeros(out.shape, dtype=torch.int32, device=features.device)
        S.check(S.lib.sis3d_roi_pool_fwd(S.ptr(features), 0, S.f32(self.spatial_scale), n, W, H, L, Cn,
                                         self.pooled_width, self.pooled_height, self.pooled_length, S.ptr(rois),
                                         S.ptr(out), S.ptr(arg), S.stream()), "roi_pool_fwd")
        self.argmax, self.rois, self.feature_size = arg, rois, features.size()
        return out

    def backward(self, grad_output):
        raise NotImplementedError("training is out of scope of the B200 inference path")
