"""CPU: lib/utils/parity.py -- what counts as 'the same integer outputs' in the bench line's parity_rate."""
import numpy as np
import torch

from lib.utils.parity import compare, scene_signature


def _P(rois, lvl, cls, keep, crops, bits=None):
    n = len(rois)
    det = np.zeros((n, 16), np.float32)
    det[:, 8] = keep
    det[:, 9:15] = crops
    P = {"rois": [torch.tensor(rois, dtype=torch.float32).reshape(n, 6)], "level_inds": [torch.tensor(lvl, dtype=torch.float32)],
         "cls_pred": torch.tensor(cls), "detections_host": det}
    if bits is not None:
        P["mask_bits"] = torch.tensor(bits, dtype=torch.uint8)
    return P


def test_compare_fields():
    rois = np.array([[0, 0, 0, 8, 8, 8], [4, 4, 4, 20, 12, 16]], np.float32)
    crops = np.array([[0, 0, 0, 8, 8, 8], [4, 4, 4, 20, 12, 16]])
    a = scene_signature(_P(rois, [1, 2], [3, 5], [1, 0], crops, bits=[1, 0, 1, 1]))
    assert compare(a, scene_signature(_P(rois + 1e-3, [1, 2], [3, 5], [1, 0], crops, bits=[1, 0, 0, 1]))) == (True, None, 0.25)
    assert compare(a, scene_signature(_P(rois[:1], [1], [3], [1], crops[:1])))[:2] == (False, "count")
    assert compare(a, scene_signature(_P(rois[::-1].copy(), [1, 2], [3, 5], [1, 0], crops)))[:2] == (False, "proposal_order")
    assert compare(a, scene_signature(_P(rois, [2, 2], [3, 5], [1, 0], crops)))[:2] == (False, "level")
    assert compare(a, scene_signature(_P(rois, [1, 2], [3, 6], [1, 0], crops)))[:2] == (False, "cls_pred")
    assert compare(a, scene_signature(_P(rois, [1, 2], [3, 5], [1, 1], crops)))[:2] == (False, "mask_keep")
    c2 = crops.copy()
    c2[1, 3] += 1
    assert compare(a, scene_signature(_P(rois, [1, 2], [3, 5], [1, 0], c2)))[:2] == (False, "crops")
