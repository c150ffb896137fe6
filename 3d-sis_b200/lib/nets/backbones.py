"""Backbone definitions (reference: lib/nets/backbones.py).

Each backbone is a *layer program*: a list of ops per stage that `Network` executes with libsis3d
kernels.  Op tuples: ("k2s2", idx, cin, cout) 2x2x2/stride-2 conv + ReLU (no bias); ("k3", idx, cin,
cout) 3x3x3 conv + ReLU (no bias); ("bneck", idx, inplanes, planes) residual bottleneck; ("pool", idx)
MaxPool3d(3,1,1).  `idx` is the position inside the reference's nn.Sequential, which fixes the
state_dict key (e.g. geometry1.2.conv1.weight) so reference checkpoints load unchanged.
"""
import torch
from torch import nn

from lib import _sis3d as S
from lib.nets.network import Act, Network
from lib.utils.config import cfg


class Base_Backbone(Network):
    def __init__(self):
        super().__init__()
        self._fc7_channels = 128
        self._net_conv_level1_channels = self._net_conv_level2_channels = self._net_conv_level3_channels = 128

    def _program(self):
        raise NotImplementedError

    def _init_backbone_classifier(self):
        self.SPEC = self._program()
        for stage in ("geometry1", "color", "geometry2"):
            if stage in self.SPEC:
                self._declare_stack(stage, self.SPEC[stage])
        p = int(cfg.CLASS_POOLING_SIZE)
        self._declare_linear("classifier.0", 256, self._net_conv_level1_channels * p ** 3)
        self._declare_linear("classifier.2", 256, 256)
        self._declare_linear("classifier.4", 128, 256)


class ScanNet_Backbone(Base_Backbone):
    """reference: backbones.py:171-231."""

    def _program(self):
        if cfg.ONLY_IMAGES:
            raise NotImplementedError("ONLY_IMAGES is not used by the released configs")
        g, c = (64, 64) if cfg.USE_IMAGES else (128, 0)
        prog = dict(geometry1=[("k2s2", 0, 2, 32), ("bneck", 2, 32, 32), ("bneck", 3, 32, 32),
                               ("k2s2", 4, 32, g), ("bneck", 6, g, 32), ("bneck", 7, g, 32)],
                    geometry2=[("k3", 0, g + c, 128), ("bneck", 2, 128, 64), ("bneck", 3, 128, 64), ("pool", 4)])
        if cfg.USE_IMAGES:
            ci = int(cfg.NUM_IMAGE_CHANNELS)
            prog["color"] = [("k2s2", 0, ci, 64), ("bneck", 2, 64, 32), ("pool", 3),
                             ("k2s2", 4, 64, c), ("bneck", 6, c, 32), ("pool", 7)]
        return prog


class SUNCG_Backbone(Base_Backbone):
    """reference: backbones.py:118-169."""

    def _program(self):
        if cfg.ONLY_IMAGES:
            raise NotImplementedError("ONLY_IMAGES is not used by the released configs")
        prog = dict(geometry1=[("k2s2", 0, 2, 64), ("bneck", 2, 64, 32), ("k2s2", 3, 64, 64), ("bneck", 5, 64, 32)],
                    geometry2=[("k3", 0, 128 if cfg.USE_IMAGES else 64, 128), ("bneck", 2, 128, 64)])
        if cfg.USE_IMAGES:
            ci = int(cfg.NUM_IMAGE_CHANNELS)
            prog["color"] = [("k2s2", 0, ci, 64), ("bneck", 2, 64, 32), ("k2s2", 3, 64, 64), ("bneck", 5, 64, 32)]
        return prog


class MaskBackbone(nn.Module):
    """Per-RoI mask head: 5 x (conv3x3x3 -> ReLU) + 1x1 -> sigmoid, no biases (reference:
    backbones.py:236-287).  Parameters live here (state_dict keys mask_backbone.geometry.N.weight); the
    kernels are launched through the owning Network so the packed weights are shared."""

    def __init__(self):
        super().__init__()
        if cfg.MASK_USE_IMAGES or cfg.MASK_ONLY_IMAGES:
            raise NotImplementedError("MASK_USE_IMAGES is off in the released configs (config.py:100)")
        from lib.nets.network import _declare
        cin = 2
        for i in (0, 2, 4, 6, 8):
            _declare(self, f"geometry.{i}.weight", (64, cin, 3, 3, 3), cin * 27)
            cin = 64
        _declare(self, "geometry.10.weight", (int(cfg.NUM_CLASSES), 64, 1, 1, 1), 64)
        self._owner = None

    def forward(self, scene, imageft=None):
        """scene [1,2,w,h,l] crop (NCDHW) -> sigmoid mask [1,num_classes,w,h,l]."""
        net = self._owner() if self._owner is not None else None
        if net is None:
            raise S.Sis3dError("mask_backbone is launched through its owning Network (which is gone)")
        net._ensure_packed()
        if scene.shape[0] != 1:
            raise S.Sis3dError("mask_backbone: batch size 1")
        scene = scene.to(next(net.parameters()).device, torch.float32).contiguous()
        w, h, l = (int(v) for v in scene.shape[2:])
        det = torch.zeros(1, 16)
        det[0, 8] = 1.0
        det[0, 12:15] = torch.tensor([w, h, l], dtype=torch.float32)
        masks = net._mask_branch(scene, det.numpy(), 1)
        return masks[0].contiguous()
