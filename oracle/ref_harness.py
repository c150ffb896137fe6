"""oracle/ref_harness.py -- TEST INFRASTRUCTURE ONLY; runs ONLY in the build container.

Imports the UNMODIFIED reference Python files from /root/reference (read-only) under a set of
import/runtime shims so that `Network.forward(blobs, 'TEST', killing_inds)` and the operator
modules execute on the CPU with torch 2.x.  Used by oracle/make_golden.py to produce the committed
fixtures under tests/golden/ that pin oracle/port.py.  Nothing on the GPU box imports this file
(/root/reference does not exist there).

Shims:
  * stub modules for deps absent from the image (easydict, ipdb, h5py, plyfile, matplotlib, skimage,
    torchnet, tensorflow, reprint, imageio) and for the two torch-0.4 FFI packages;
  * Tensor.cuda / Module.cuda / torch.cuda.* -> no-ops (routes nms -> numpy cpu_nms and RoI pooling
    -> the reference's CPU C kernel, compiled unmodified into oracle/_ref/libref_roi_cpu.so);
  * int64 `/` -> floor division (torch-0.4 LongTensor semantics, projection.py:68,70,80,82);
  * legacy instance-style autograd.Function made callable (roi_pool.py:9-49);
  * yaml.load default Loader; cwd = reference root (anchor tables are opened by relative path).
"""
from __future__ import annotations

import ctypes
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SIS3D_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


class _EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = __setitem__


class _THFloat(ctypes.Structure):
    _fields_ = [("data", ctypes.POINTER(ctypes.c_float)), ("size", ctypes.c_long * 8)]


def _th(t):
    s = _THFloat()
    s.data = ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))
    for i, v in enumerate(t.shape):
        s.size[i] = v
    return s


_installed = False
EXT_CALLS = {"gpu_nms": 0, "roi_pooling_forward_cuda": 0}  # how often the rebound native entry points were used (gpu mode)


def _libsis3d_ext_stubs():
    """INTEGRATION.md section 2, executed: ctypes stubs with the signatures of the reference's two cffi extension modules
    (`gpu_nms`: lib/layer_utils/nms/src/nms_cuda.h:1; `roi_pooling_forward_cuda`: lib/layer_utils/roi_pooling/src/
    roi_pooling_cuda.h:1-2) over the C ABI of libsis3d.so."""
    lib = ctypes.CDLL(os.path.join(os.path.dirname(_HERE), "3d-sis_b200", "lib", "libsis3d.so"))
    lib.sis3d_nms_workspace_bytes.restype = ctypes.c_size_t
    lib.sis3d_strerror.restype = ctypes.c_char_p

    def gpu_nms(keep, num_out, boxes, thresh):  # keep / num_out: CPU LongTensors, boxes: CUDA [N,6] sorted by score
        EXT_CALLS["gpu_nms"] += 1
        n = boxes.size(0)
        boxes = boxes.contiguous().float()
        ws = torch.empty(max(int(lib.sis3d_nms_workspace_bytes(n)), 8), dtype=torch.uint8, device=boxes.device)
        keep_d = torch.empty(max(n, 1), dtype=torch.int64, device=boxes.device)
        num_d = torch.zeros(1, dtype=torch.int32, device=boxes.device)
        rc = lib.sis3d_nms(ctypes.c_void_p(boxes.data_ptr()), n, ctypes.c_float(thresh), ctypes.c_void_p(keep_d.data_ptr()),
                           ctypes.c_void_p(num_d.data_ptr()), ctypes.c_void_p(ws.data_ptr()),
                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc:
            raise RuntimeError(lib.sis3d_strerror(rc).decode())
        k = int(num_d.item())
        num_out[0] = k
        keep[:k] = keep_d[:k].cpu()
        return 1

    def roi_pooling_forward_cuda(pw, ph, pl, scale, features, rois, output, argmax):
        EXT_CALLS["roi_pooling_forward_cuda"] += 1
        _, C, W, H, L = features.size()
        features, rois = features.contiguous(), rois.contiguous().float()
        rc = lib.sis3d_roi_pool_fwd(ctypes.c_void_p(features.data_ptr()), 0, ctypes.c_float(scale), rois.size(0), W, H, L, C,
                                    int(pw), int(ph), int(pl), ctypes.c_void_p(rois.data_ptr()),
                                    ctypes.c_void_p(output.data_ptr()), ctypes.c_void_p(argmax.data_ptr()),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        return 1 if rc == 0 else 0

    return gpu_nms, roi_pooling_forward_cuda


def install(gpu=False):
    """Install all shims and put the reference root on sys.path.  Idempotent.
    gpu=False: everything on the CPU (`.cuda()` no-ops; numpy NMS, the reference's CPU RoI-pooling C kernel).
    gpu=True : the reference's Python runs on the GPU as written (cuDNN convs in fp32) and its two native extensions are the
               ctypes stubs of INTEGRATION.md section 2 bound to libsis3d.so -- the drop-in test of the operator boundary."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(os.path.join(REF, "lib")):
        raise RuntimeError(f"reference tree not found at {REF}")

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    stub("easydict", EasyDict=_EasyDict)
    stub("ipdb", set_trace=lambda *a, **k: None)
    for name in ("h5py", "plyfile", "torchnet", "tensorflow", "reprint", "imageio", "tqdm"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                stub(name, PlyData=object, PlyElement=object, output=object, tqdm=lambda x, **k: x)
    mpl = stub("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = stub("matplotlib.pyplot")
    sk = stub("skimage")
    sk.transform = stub("skimage.transform")

    # FFI packages ------------------------------------------------------------------------------
    ref_roi = ctypes.CDLL(os.path.join(_HERE, "_ref", "libref_roi_cpu.so"))

    def roi_pooling_forward(pw, ph, pl, scale, features, rois, output):
        f, r, o = features.contiguous(), rois.contiguous().float(), output
        return ref_roi.roi_pooling_forward(int(pw), int(ph), int(pl), ctypes.c_float(scale),
                                           ctypes.byref(_th(f)), ctypes.byref(_th(r)), ctypes.byref(_th(o)))

    ext_roi = stub("lib.layer_utils.roi_pooling._ext")
    ext_nms = stub("lib.layer_utils.nms._ext")
    if gpu:
        gpu_nms, roi_fwd_cuda = _libsis3d_ext_stubs()
        ext_roi.roi_pooling = stub("lib.layer_utils.roi_pooling._ext.roi_pooling", roi_pooling_forward=roi_pooling_forward,
                                   roi_pooling_forward_cuda=roi_fwd_cuda)
        ext_nms.nms = stub("lib.layer_utils.nms._ext.nms", gpu_nms=gpu_nms)
        # torch-0.4 semantics the reference relies on: np.where(<cuda tensor>) copies to the host implicitly
        # (proposal_layer.py:36-43); fp32 convolutions (no TF32 inside cuDNN / cuBLAS)
        _arr = torch.Tensor.__array__
        torch.Tensor.__array__ = lambda self, *a, **k: _arr(self.detach().cpu() if self.is_cuda else self, *a, **k)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    else:
        ext_roi.roi_pooling = stub("lib.layer_utils.roi_pooling._ext.roi_pooling", roi_pooling_forward=roi_pooling_forward)
        ext_nms.nms = stub("lib.layer_utils.nms._ext.nms")
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.synchronize = lambda *a, **k: None
        torch.cuda.empty_cache = lambda *a, **k: None

    # runtime patches ----------------------------------------------------------------------------
    _truediv = torch.Tensor.__truediv__

    def _div(self, other):
        if not self.is_floating_point() and isinstance(other, (int, np.integer)):
            return torch.div(self, other, rounding_mode="floor")
        if not self.is_floating_point() and isinstance(other, torch.Tensor) and not other.is_floating_point():
            return torch.div(self, other, rounding_mode="floor")
        return _truediv(self, other)

    torch.Tensor.__truediv__ = _div
    import yaml
    _load = yaml.load
    yaml.load = lambda stream, Loader=None: _load(stream, Loader=Loader or yaml.SafeLoader)

    sys.path.insert(0, REF)
    os.chdir(REF)
    import lib.nets.backbones  # noqa: F401  (must precede lib.nets.network: circular import)
    import lib.layer_utils.roi_pooling.roi_pool as rp

    _Fn = rp.RoIPoolFunction

    class _CallableRoIPool:
        """Instance-style wrapper executing the reference's own forward body."""

        def __init__(self, pw, ph, pl, scale):
            self.pooled_height, self.pooled_width, self.pooled_length = int(ph), int(pw), int(pl)
            self.spatial_scale = float(scale)
            self.argmax = self.rois = self.feature_size = None

        def __call__(self, features, rois):
            return _Fn.forward(self, features, rois)

    rp.RoIPoolFunction = _CallableRoIPool
    import lib.nets.network as net_mod
    net_mod.RoIPoolFunction = _CallableRoIPool
    _installed = True


def load_cfg(yml_rel, gpu=False, **over):
    """cfg_from_file + NUM_CLASSES as derived by main.py:44-50."""
    install(gpu)
    from lib.utils.config import cfg, cfg_from_file
    cfg_from_file(os.path.join(REF, "experiments", "cfgs", yml_rel))
    cfg.NUM_CLASSES = 26 if 'SUNCG' in yml_rel else 19  # = #labels with weight>0 in cfg.LABEL_MAP (main.py:44-50)
    for k, v in over.items():
        cfg[k] = v
    return cfg


def build_net(cfg, weights):
    """getattr(backbones, cfg.NET)() + init_modules + load_state_dict (trainval.py:88-91)."""
    from lib.nets import backbones
    net = getattr(backbones, cfg.NET)()
    net.init_modules()
    sd = {k: torch.from_numpy(np.array(v)) for k, v in weights.items()}
    missing = net.load_state_dict(sd, strict=True)
    net.eval()
    return net


def reference_forward(net, cfg, data, views):
    """One scene through the UNMODIFIED reference exactly as its driver does it (lib/model/trainval.py:797-822 with the
    MAX_VOLUME=0 'CPU path' semantics): per-view ProjectionHelper.compute_projection, killing_inds for views without a valid
    projection, index lists stacked densely, then Network.forward(blobs, 'TEST', killing_inds).  Returns net._predictions."""
    from lib.layer_utils.projection import ProjectionHelper
    blobs = {"data": torch.from_numpy(np.ascontiguousarray(data)), "id": ["synthetic"], "gt_box": [torch.zeros(0, 7)],
             "gt_mask": [[]]}
    killing = None
    if views is not None:
        helper = ProjectionHelper(cfg.INTRINSIC, cfg.PROJ_DEPTH_MIN, cfg.PROJ_DEPTH_MAX, cfg.DEPTH_SHAPE,
                                  blobs["data"].shape[-3:], cfg.VOXEL_SIZE)
        w2g = torch.from_numpy(views["world2grid"])
        vol = int(np.prod(blobs["data"].shape[-3:]))
        if vol > cfg.MAX_VOLUME or len(views["depths"]) > cfg.MAX_IMAGE:  # trainval.py:800-801: projection on the CPU
            dev = lambda t: t
        else:                                                              # :802-803 (`.cuda()` is a no-op under the CPU shims)
            dev = lambda t: t.cuda()
        maps = [helper.compute_projection(dev(torch.from_numpy(d)), dev(torch.from_numpy(c)), dev(w2g))
                for d, c in zip(views["depths"], views["poses"])]
        killing = [i for i, m in enumerate(maps) if m is None]
        real = [m for m in maps if m is not None]
        blobs["proj_ind_3d"] = [torch.stack([m[0] for m in real])]
        blobs["proj_ind_2d"] = [torch.stack([m[1] for m in real])]
        blobs["nearest_images"] = {"images": [torch.from_numpy(views["feats"])]}
    net.forward(blobs, "TEST", killing)
    return net._predictions
