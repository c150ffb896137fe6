"""Pins oracle/port.py against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import sis3d_synth as synth
from conftest import load_golden

from sis3d_synth import CASES, build_case  # noqa: E402,F401  (shared with smoke() and bench.py)


def sub(t, step):
    return np.asarray(t, dtype=np.float32).reshape(-1)[::step]


@pytest.mark.parametrize("tag", list(CASES))
def test_port_matches_reference_forward(oracle, tag):
    c = CASES[tag]
    if tag == "cfg2_96x48x96":
        torch.set_num_threads(max(1, torch.get_num_threads()))
    g = load_golden(f"forward_{tag}.npz")
    cfg, w, data, views = build_case(oracle, c)
    out = oracle.forward(cfg, w, data, views)
    if c["use_images"]:
        assert list(g["killing_inds"]) == out["killing_inds"]
        for i in range(c["n_img"]):
            m = oracle.compute_projection(cfg, views["depths"][i], views["poses"][i], views["world2grid"], c["dims"])
            assert np.array_equal(m[0], g[f"proj3d_{i}"]), f"view {i} lin3d"
            assert np.array_equal(m[1], g[f"proj2d_{i}"]), f"view {i} lin2d"
        assert np.array_equal(sub(out["imageft"], 97), g["imageft_sub"])  # exact: pure copy/max
    for lvl in (1, 2):
        np.testing.assert_allclose(sub(out[f"rpn_prob_level{lvl}"], 5), g[f"rpn_prob_sub_l{lvl}"], atol=2e-6)
        np.testing.assert_allclose(sub(out[f"rpn_deltas_level{lvl}"], 11), g[f"rpn_bbox_sub_l{lvl}"], atol=2e-5)
    assert out["rois"].shape == g["rois"].shape
    np.testing.assert_allclose(out["rois"].numpy(), g["rois"], atol=1e-3)
    np.testing.assert_allclose(out["roi_scores"].numpy().reshape(-1), g["roi_scores"].reshape(-1), atol=2e-6)
    assert np.array_equal(out["level_inds"].numpy(), g["level_inds"])
    np.testing.assert_allclose(out["cls_prob"].numpy(), g["cls_prob"], atol=1e-5)
    assert np.array_equal(out["cls_pred"].numpy(), g["cls_pred"])
    np.testing.assert_allclose(out["bbox_pred"].numpy(), g["bbox_pred"], atol=1e-5)
    if c["use_mask"]:
        np.testing.assert_allclose(out["pred_box"], g["pred_box"], atol=1e-3)
        assert np.array_equal(out["mask_keep"], g["mask_keep"])
        kept = np.nonzero(out["mask_keep"])[0]
        for j, i in enumerate(kept):
            m = out["mask_pred"][j][0].numpy()
            np.testing.assert_allclose(m[int(g["cls_pred"][i])], g[f"mask_{j}_cls"], atol=1e-5)
            np.testing.assert_allclose(m.reshape(-1)[::13], g[f"mask_{j}_allcls_sub"], atol=1e-5)


WHOLE_SCENES = {  # BASELINE configs[2] shapes; seeds shared with tests/test_gpu_forward.py's whole-scene tests
    "scene_88x44x88": dict(cfgname="scannet", dims=(88, 44, 88), n_img=8, seed=505, use_images=True, use_mask=True),
    "stress_208x48x160": dict(cfgname="scannet", dims=(208, 48, 160), n_img=40, seed=606, use_images=True, use_mask=True),
}


@pytest.mark.parametrize("tag", list(WHOLE_SCENES))
def test_port_matches_reference_whole_scene(oracle, tag):
    """The oracle at the fully-convolutional whole-scene shapes (up to 1.6 M voxels, 40 views), against compact fixtures of
    the unmodified reference (oracle/make_golden.py scene stress): per-view index lists by sha1, strided samples of the
    dense tensors, the complete RoI / class / box tables and strided masks."""
    import hashlib
    c = WHOLE_SCENES[tag]
    g = load_golden(f"forward_{tag}.npz")
    cfg, w, data, views = build_case(oracle, c)
    out = oracle.forward(cfg, w, data, views)
    assert list(g["killing_inds"]) == out["killing_inds"]
    sha = lambda a: np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
    for i in range(c["n_img"]):
        m = oracle.compute_projection(cfg, views["depths"][i], views["poses"][i], views["world2grid"], c["dims"])
        assert len(m[0]) == int(g[f"proj_count_{i}"]), f"view {i} count"
        assert np.array_equal(sha(m[0].astype(np.int32)), g[f"proj3d_sha_{i}"]), f"view {i} lin3d"
        assert np.array_equal(sha(m[1].astype(np.int16)), g[f"proj2d_sha_{i}"]), f"view {i} lin2d"
    assert np.array_equal(sub(out["imageft"], 997), g["imageft_sub"])
    for lvl in (1, 2):
        np.testing.assert_allclose(sub(out[f"rpn_prob_level{lvl}"], 53), g[f"rpn_prob_sub_l{lvl}"], atol=2e-6)
        np.testing.assert_allclose(sub(out[f"rpn_deltas_level{lvl}"], 101), g[f"rpn_bbox_sub_l{lvl}"], atol=2e-5)
    assert out["rois"].shape == g["rois"].shape
    np.testing.assert_allclose(out["rois"].numpy(), g["rois"], atol=1e-3)
    np.testing.assert_allclose(out["roi_scores"].numpy().reshape(-1), g["roi_scores"].reshape(-1), atol=2e-6)
    assert np.array_equal(out["level_inds"].numpy(), g["level_inds"])
    np.testing.assert_allclose(out["cls_prob"].numpy(), g["cls_prob"], atol=1e-5)
    assert np.array_equal(out["cls_pred"].numpy(), g["cls_pred"])
    np.testing.assert_allclose(out["bbox_pred"].numpy(), g["bbox_pred"], atol=1e-5)
    np.testing.assert_allclose(out["pred_box"], g["pred_box"], atol=1e-3)
    assert np.array_equal(out["mask_keep"], g["mask_keep"])
    for j, i in enumerate(np.nonzero(out["mask_keep"])[0]):
        m = out["mask_pred"][j][0].numpy()[int(g["cls_pred"][i])]
        np.testing.assert_allclose(m.reshape(-1)[::7], g[f"mask_{j}_cls_sub"], atol=1e-5)
        assert abs(int((m > 0.5).sum()) - int(g[f"mask_{j}_cls_on"])) <= 1  # a voxel within 1e-5 of the threshold may flip


def test_port_operators(oracle):
    g = load_golden("operators.npz")
    for seed in range(6):
        s, n, thr = g[f"nms_cfg_{seed}"]
        b = synth.make_nms_boxes(int(s), int(n))
        keep = oracle.nms3d(b, float(thr), fma_mode=0)
        assert np.array_equal(keep, g[f"nms_keep_{seed}"]), f"nms case {seed}"
    rng = np.random.default_rng(11)
    feat = rng.standard_normal((1, 16, 24, 12, 24)).astype(np.float32)
    out, arg = oracle.roi_pool3d(feat, g["roi_rois"], (4, 4, 4), 0.25)
    assert np.array_equal(out, g["roi_out"])
    # argmax consistency: value at argmax equals the pooled value (or bin empty -> -1 / 0)
    flat = feat.reshape(-1)
    ok = np.where(arg >= 0, flat[np.maximum(arg, 0)], 0.0)
    assert np.array_equal(ok, out)
    a1 = oracle.generate_anchors([5, 3, 4], oracle.read_anchor_table("scannet14_3.txt"))
    a2 = oracle.generate_anchors([2, 3, 2], oracle.read_anchor_table("scannet14_11.txt"))
    assert np.array_equal(a1, g["anchors_l1_5x3x4"]) and np.array_equal(a2, g["anchors_l2_2x3x2"])
    pb = oracle.bbox_transform_inv(a1, g["decode_deltas"])
    assert np.array_equal(pb.numpy(), g["decode_boxes"])
    assert np.array_equal(oracle.clip_boxes(pb, [20, 12, 16]).numpy(), g["decode_clipped"])
