"""CPU-only: the C-ABI library loads and exports every symbol include/sis3d.h declares."""
import os
import re

from conftest import ROOT


def test_header_symbols_exported():
    from lib import _sis3d as S
    hdr = open(os.path.join(ROOT, "include", "sis3d.h")).read()
    declared = set(re.findall(r"\b(sis3d_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    missing = [s for s in sorted(declared) if not hasattr(S.lib, s)]
    assert not missing, f"declared in sis3d.h but not exported by libsis3d.so: {missing}"
    assert set(S.SYMBOLS) <= declared
    assert S.lib.sis3d_version() >= 100
    assert S.lib.sis3d_strerror(-3).decode() == "workspace too small"


def test_region_struct_layout():
    from lib import _sis3d as S
    assert S.REGION_BYTES == 104  # 3*8 + 3*4 + 3*4 + 3*8 + 3*8 + 4 + 4


def test_no_cpu_fallback():
    import pytest
    import torch
    from lib import _sis3d as S
    from lib.layer_utils.nms_wrapper import nms
    with pytest.raises(S.Sis3dError):
        nms(torch.zeros(4, 6), 0.5)


def test_state_dict_contract():
    """Parameter names/shapes equal the reference's state_dict."""
    import sis3d_synth as synth
    from lib.utils.config import cfg, cfg_from_file, cfg_reset
    cfg_reset()
    cfg_from_file(os.path.join(ROOT, "3d-sis_b200", "experiments", "cfgs", "ScanNet", "rpn_class_mask_5.yml"))
    cfg.NUM_CLASSES, cfg.USE_IMAGES_GT = 19, True
    from lib.nets import backbones
    net = backbones.ScanNet_Backbone()
    net.init_modules()
    want = synth.param_shapes()
    got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert got == dict(want)
