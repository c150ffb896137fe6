"""oracle/decision_margins.py -- TEST INFRASTRUCTURE ONLY (CPU; uses the oracle, never the product path).

    python oracle/decision_margins.py > profiles/r2_decision_margins.txt

How close the integer decisions of the detector sit to their thresholds on the synthetic bench scenes: the gaps between
neighbouring pre-NMS scores (lib/layer_utils/proposal_layer.py:150-167 in the reference sorts them), the distance of the
decisive NMS IoUs from RPN_NMS_THRESH (:170-176), the class-argmax and CLASS_THRESH margins (lib/nets/network.py:296-301) and
the distance of the predicted box corners from the .5 rounding boundary of the mask crop.  These are the margins a conv math
mode's error has to stay under for `parity_rate` (lib/utils/parity.py) to report a scene as exact.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-sis_b200"))
sys.path.insert(0, ROOT)
from oracle import port  # noqa: E402
import sis3d_synth as synth  # noqa: E402

ULP = 2.0 ** -24  # spacing of fp32 just below 1.0


def margins(seed, dims, n_img):
    c = dict(cfgname="scannet", dims=dims, n_img=n_img, seed=seed, use_images=True, use_mask=True)
    cfg, w, data, views = synth.build_case(port, c)
    o = port.forward(cfg, w, data, views)
    s = o["rpn_all_scores"][o["rpn_order"]].astype(np.float64)
    g = -np.diff(s)
    keep = np.asarray(o["nms_keep"])
    iou = port.iou_matrix(o["rpn_sorted_boxes"], 0)
    m = np.abs(iou[keep] - np.float32(cfg.RPN_NMS_THRESH))
    for r, i in enumerate(keep):
        m[r, :i + 1] = np.inf  # only (kept box, later candidate) pairs decide anything
    cp = np.sort(o["cls_prob"].numpy(), 1)
    pb = np.asarray(o["pred_box"])
    return dict(seed=seed, dims="x".join(map(str, dims)), candidates=len(s), score_lo=float(s[-1]), score_hi=float(s[0]),
                ties=int((g == 0).sum()), one_ulp=int(((g > 0) & (g <= 2 * ULP)).sum()),
                below_1e6=int(((g > 2 * ULP) & (g <= 1e-6)).sum()), iou=float(m.min()),
                cls=float((cp[:, -1] - cp[:, -2]).min()), conf=float(np.abs(np.asarray(o["pred_conf"]) - cfg.CLASS_THRESH).min()),
                crop=float(np.abs((pb - np.floor(pb)) - 0.5).min()))


if __name__ == "__main__":
    rows = [margins(s, (208, 48, 160), 40) for s in (3000, 3001, 3002, 3003)]      # bench.py --config cfg3, rank 0
    rows += [margins(s, (96, 48, 96), 5) for s in range(1000, 1008)]               # bench.py --config cfg2, first 8 chunks
    print("pre-NMS candidates (top RPN_PRE_NMS_TOP_N by score) and decision margins, oracle fp32 (CPU)")
    print(f"{'seed':>5} {'dims':>11} {'cand':>5} {'score range':>21} {'ties':>5} {'<=1ulp':>7} {'<=1e-6':>7} "
          f"{'|IoU-thr|':>10} {'cls gap':>9} {'|conf-thr|':>10} {'crop .5':>9}")
    for r in rows:
        print(f"{r['seed']:>5} {r['dims']:>11} {r['candidates']:>5} {r['score_lo']:.6f} .. {r['score_hi']:.6f} {r['ties']:>5} "
              f"{r['one_ulp']:>7} {r['below_1e6']:>7} {r['iou']:>10.2e} {r['cls']:>9.2e} {r['conf']:>10.2e} {r['crop']:>9.2e}")
    print("\nties / <=1ulp / <=1e-6: neighbouring candidates in the sorted list whose scores are equal / one fp32 step apart / "
          "closer than 1e-6.\nThe candidates' scores occupy ~1.2e-3 just below 1.0 (fp32 step 6e-8), so 15-25 neighbouring pairs per "
          "scene are decided by the last bit\nof a softmax output; every other decision has a margin of 1e-5 or more.")
