"""Exact-match rate of the detector's integer outputs between two conv-math modes of the product path.

North star: "bit-exact for NMS indices / RoI bins".  The integer outputs of a scene are the proposal count and order (the
NMS keep list: rows must be the same anchors in the same score order), the pyramid level ids, the class argmax, the
mask-keep flags and the integer crop bounds of the mask head (lib/layer_utils/proposal_layer.py:181-197,
lib/nets/network.py:296-301 in the reference).  `scene_signature` extracts them from `Network._predictions`;
`parity_rate` runs the same scenes through a candidate math mode and through the fp32 CUDA-core mode (which the GPU tests
pin bit-exact to the unmodified reference on the golden cases) and reports the fraction of scenes where ALL of them agree.
"""
import numpy as np


def scene_signature(P):
    rois = P["rois"][0].detach().cpu().numpy()
    sig = dict(n=int(rois.shape[0]), rois=rois, level=P["level_inds"][0].detach().cpu().numpy().astype(np.int64))
    if "cls_pred" in P:
        sig["cls_pred"] = P["cls_pred"].detach().cpu().numpy().astype(np.int64)
    if "detections_host" in P:
        det = np.asarray(P["detections_host"])
        sig["mask_keep"] = det[:, 8] > 0.5
        sig["crops"] = det[:, 9:15].astype(np.int64)
    if "mask_bits" in P:
        sig["mask_bits"] = P["mask_bits"].detach().cpu().numpy().copy()
    return sig


def compare(a, b, box_tol=0.05):
    """-> (exact: bool, first differing field or None, fraction of thresholded mask voxels that differ or None)."""
    if a["n"] != b["n"]:
        return False, "count", None
    if a["n"] and float(np.abs(a["rois"] - b["rois"]).max()) >= box_tol:
        return False, "proposal_order", None
    for k in ("level", "cls_pred", "mask_keep", "crops"):
        if k in a and not np.array_equal(a[k], b[k]):
            return False, k, None
    flips = None
    if "mask_bits" in a and "mask_bits" in b and a["mask_bits"].shape == b["mask_bits"].shape and a["mask_bits"].size:
        flips = float((a["mask_bits"] != b["mask_bits"]).mean())
    return True, None, flips


def parity_rate(make_net, blobs_list, mode, ref_mode="fp32", ref_sigs=None):
    """make_net(mode) -> Network.  Returns (summary dict, reference signatures for reuse)."""
    import torch
    if ref_sigs is None:
        net = make_net(ref_mode)
        ref_sigs = []
        for b in blobs_list:
            ref_sigs.append(scene_signature(net.forward(b, "TEST", None)))
        del net
        torch.cuda.synchronize()
    net = make_net(mode)
    exact, why, flips = 0, {}, []
    for b, r in zip(blobs_list, ref_sigs):
        ok, field, fl = compare(scene_signature(net.forward(b, "TEST", None)), r)
        exact += int(ok)
        if not ok:
            why[field] = why.get(field, 0) + 1
        if fl is not None:
            flips.append(fl)
    del net
    torch.cuda.synchronize()
    return dict(mode=mode, against=ref_mode, scenes=len(blobs_list), exact_scenes=exact, rate=exact / max(1, len(blobs_list)),
                first_mismatch_fields=why,
                thresholded_mask_voxel_flip_fraction=float(np.mean(flips)) if flips else None), ref_sigs
