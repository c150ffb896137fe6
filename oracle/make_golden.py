"""oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY; run in the build container:

    python oracle/make_golden.py            # writes tests/golden/*.npz

Executes the UNMODIFIED reference (/root/reference) through oracle/ref_harness.py on the seeded
synthetic inputs of 3d-sis_b200/sis3d_synth.py and stores compact known-answer fixtures.  The
fixtures pin oracle/port.py (tests/test_oracle_golden.py) and are a second, reference-generated
target for the CUDA parity tests.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-sis_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sis3d_synth as synth  # noqa: E402
import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sub(t, step=7):
    """Deterministic strided subsample + moments of a big tensor."""
    a = np.asarray(t, dtype=np.float32).reshape(-1)
    return a[::step].copy(), np.array([a.astype(np.float64).sum(), np.abs(a).astype(np.float64).sum(),
                                       float(a.max()), float(a.min())])


def digest(a):
    """sha1 of an index array's bytes (whole-scene cases: 40 views of up to 83 k indices would be 20 MB as arrays)."""
    import hashlib
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8).copy()


def run_forward(tag, yml, dims, n_img, seed, use_images, use_mask, net_name, compact=False):
    cfg = rh.load_cfg(yml, USE_IMAGES=use_images, USE_IMAGES_GT=True, USE_MASK=use_mask)
    from lib.layer_utils.projection import ProjectionHelper
    num_classes = cfg.NUM_CLASSES
    a2 = cfg.NUM_ANCHORS_LEVEL2
    w = synth.make_weights(seed=0, net=net_name, use_images=use_images, num_classes=num_classes,
                           a1=cfg.NUM_ANCHORS_LEVEL1, a2=a2, use_mask=use_mask)
    net = rh.build_net(cfg, w)
    data, boxes = synth.make_scene(seed, dims)
    blobs = {"data": torch.from_numpy(data), "id": [f"synthetic_{tag}"], "gt_box": [torch.zeros(0, 7)],
             "gt_mask": [[]]}
    g = {"dims": np.array(dims), "seed": np.array(seed), "n_img": np.array(n_img)}
    killing = None
    if use_images:
        v = synth.make_views(seed, dims, n_img, boxes,
                             intrinsic=np.array(cfg.INTRINSIC, dtype=np.float32))
        helper = ProjectionHelper(cfg.INTRINSIC, cfg.PROJ_DEPTH_MIN, cfg.PROJ_DEPTH_MAX, cfg.DEPTH_SHAPE,
                                  blobs["data"].shape[-3:], cfg.VOXEL_SIZE)
        # trainval.py:797-820 with MAX_VOLUME=0 semantics (projection on CPU)
        maps = [helper.compute_projection(torch.from_numpy(d), torch.from_numpy(c), torch.from_numpy(v["world2grid"]))
                for d, c in zip(v["depths"], v["poses"])]
        killing = [i for i, m in enumerate(maps) if m is None]
        real = [m for m in maps if m is not None]
        for i, m in enumerate(maps):
            if m is not None:
                k = int(m[0][0])
                p3, p2 = m[0][1:1 + k].numpy().astype(np.int32), m[1][1:1 + k].numpy().astype(np.int16)
                if compact:
                    g[f"proj_count_{i}"], g[f"proj3d_sha_{i}"], g[f"proj2d_sha_{i}"] = np.array(k), digest(p3), digest(p2)
                else:
                    g[f"proj3d_{i}"], g[f"proj2d_{i}"] = p3, p2
        g["killing_inds"] = np.array(killing, dtype=np.int64)
        blobs["proj_ind_3d"] = [torch.stack([m[0] for m in real])]
        blobs["proj_ind_2d"] = [torch.stack([m[1] for m in real])]
        blobs["nearest_images"] = {"images": [torch.from_numpy(v["feats"])]}
    net.forward(blobs, "TEST", killing)
    P = net._predictions
    if use_images:
        g["imageft_sub"], g["imageft_stats"] = sub(net._imageft, 997 if compact else 97)
    g["rois"] = P["rois"][0].numpy()
    g["roi_scores"] = P["roi_scores"][0].numpy()
    g["level_inds"] = P["level_inds"][0].numpy()
    for lvl in (1, 2):
        g[f"rpn_prob_sub_l{lvl}"], g[f"rpn_prob_stats_l{lvl}"] = sub(P[f"rpn_cls_prob_level{lvl}"][0, 1], 53 if compact else 5)
        g[f"rpn_bbox_sub_l{lvl}"], g[f"rpn_bbox_stats_l{lvl}"] = sub(P[f"rpn_bbox_pred_level{lvl}"], 101 if compact else 11)
    g["cls_prob"] = P["cls_prob"].numpy()
    g["cls_pred"] = P["cls_pred"].numpy()
    g["bbox_pred"] = P["bbox_pred"].numpy()
    if use_mask:
        # the class-specific decode the driver performs (trainval.py:825-858)
        from lib.utils.bbox_transform import bbox_transform_inv, clip_boxes
        pc = g["cls_pred"]
        reg = np.stack([g["bbox_pred"][i, pc[i] * 6:(pc[i] + 1) * 6] for i in range(len(pc))]) if len(pc) else np.zeros((0, 6))
        pb = clip_boxes(bbox_transform_inv(P["rois"][0], torch.from_numpy(reg).float()), net._scene_info[:3]).numpy()
        g["pred_box"] = pb
        masks = P["mask_pred"][0]
        conf = np.array([g["cls_prob"][i, pc[i]] for i in range(len(pc))])
        keep = conf > cfg.CLASS_THRESH
        for i, b in enumerate(pb):
            if round(b[0]) >= round(b[3]) or round(b[1]) >= round(b[4]) or round(b[2]) >= round(b[5]):
                keep[i] = False
        g["mask_keep"] = keep
        assert keep.sum() == len(masks), (keep.sum(), len(masks))
        for j, i in enumerate(np.nonzero(keep)[0]):
            m = masks[j][0].numpy()  # [num_classes,w,h,l]
            if compact:  # the predicted class's mask, strided, + the number of voxels above the driver's 0.5 threshold
                g[f"mask_{j}_cls_sub"] = m[pc[i]].reshape(-1)[::7].copy()
                g[f"mask_{j}_cls_on"] = np.array(int((m[pc[i]] > 0.5).sum()))
            else:
                g[f"mask_{j}_cls"] = m[pc[i]]
                g[f"mask_{j}_allcls_sub"] = m.reshape(-1)[::13].copy()
    path = os.path.join(OUT, f"forward_{tag}.npz")
    np.savez_compressed(path, **g)
    print(f"[golden] {tag}: rois={len(g['rois'])} masks={int(g.get('mask_keep', np.zeros(0)).sum())} "
          f"killed={killing} -> {os.path.getsize(path) / 1e3:.0f} kB")


def run_operators():
    rh.install()
    from lib.layer_utils.nms.pth_nms import cpu_nms
    from lib.layer_utils.roi_pooling.roi_pool import RoIPoolFunction
    from lib.utils.bbox_transform import bbox_transform_inv, clip_boxes
    g = {}
    for seed, n, thr in ((0, 400, 0.1), (1, 400, 0.35), (2, 1000, 0.5), (3, 64, 0.1), (4, 65, 0.7), (5, 1, 0.1)):
        b = synth.make_nms_boxes(seed, n)
        g[f"nms_keep_{seed}"] = np.asarray(cpu_nms(b, thr), dtype=np.int64)
        g[f"nms_cfg_{seed}"] = np.array([seed, n, thr])
    rng = np.random.default_rng(11)
    feat = rng.standard_normal((1, 16, 24, 12, 24)).astype(np.float32)
    rois = synth.make_nms_boxes(21, 40)
    rois[0] = [5, 5, 5, 5, 5, 5]            # degenerate -> forced 1x1x1
    rois[1] = [90, 40, 90, 96, 48, 96]      # touches the far border
    rois[2] = [0, 0, 0, 96, 48, 96]         # whole volume
    rois[3] = [10.5, 3.25, 7.75, 11.0, 3.5, 8.0]
    out = RoIPoolFunction(4, 4, 4, 0.25)(torch.from_numpy(feat), torch.from_numpy(rois))
    g["roi_feat_seed"], g["roi_rois"], g["roi_out"] = np.array(11), rois, out.numpy()
    from lib.utils.config import cfg
    from lib.layer_utils.generate_anchors import generate_anchors
    cfg.NUM_ANCHORS_LEVEL1, cfg.NUM_ANCHORS_LEVEL2, cfg.NUM_ANCHORS_LEVEL3 = 3, 11, 0
    cfg.ANCHORS_TYPE_LEVEL1, cfg.ANCHORS_TYPE_LEVEL2 = "scannet14_3.txt", "scannet14_11.txt"
    a1, a2, _ = generate_anchors([5, 3, 4], [2, 3, 2], [], [4, 4, 4])
    g["anchors_l1_5x3x4"], g["anchors_l2_2x3x2"] = a1, a2
    d = rng.normal(0, 0.3, (a1.shape[0], 6)).astype(np.float32)
    pb = bbox_transform_inv(torch.from_numpy(a1), torch.from_numpy(d))
    g["decode_deltas"], g["decode_boxes"] = d, pb.numpy()
    g["decode_clipped"] = clip_boxes(pb, [20, 12, 16]).numpy()
    np.savez_compressed(os.path.join(OUT, "operators.npz"), **g)
    print("[golden] operators written")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    which = sys.argv[1:] or ["ops", "cfg1", "odd", "cfg2", "suncg", "scene", "stress"]
    if "ops" in which:
        run_operators()
    if "cfg1" in which:
        run_forward("cfg1_32", "ScanNet/rpn_class_mask_5.yml", (32, 32, 32), 0, 101, False, False, "ScanNet_Backbone")
    if "odd" in which:
        run_forward("odd_45x27x41", "ScanNet/rpn_class_mask_5.yml", (45, 27, 41), 3, 202, True, True, "ScanNet_Backbone")
    if "cfg2" in which:
        run_forward("cfg2_96x48x96", "ScanNet/rpn_class_mask_5.yml", (96, 48, 96), 5, 303, True, True, "ScanNet_Backbone")
    if "suncg" in which:
        run_forward("suncg_40x24x40", "SUNCG/rpn_class_mask_5.yml", (40, 24, 40), 3, 404, True, True, "SUNCG_Backbone")
    # whole-scene shapes of BASELINE configs[2] (same seeds as tests/test_gpu_forward.py's whole-scene tests), compact fixtures
    if "scene" in which:
        run_forward("scene_88x44x88", "ScanNet/rpn_class_mask_5.yml", (88, 44, 88), 8, 505, True, True, "ScanNet_Backbone",
                    compact=True)
    if "stress" in which:
        run_forward("stress_208x48x160", "ScanNet/rpn_class_mask_5.yml", (208, 48, 160), 40, 606, True, True,
                    "ScanNet_Backbone", compact=True)
