#!/usr/bin/env python
"""Flake hunt for the scene loop: forward_pipelined must yield exactly what the synchronous forward yields.
  python tools/loop_pipelined.py [--nets 8] [--reps 25] [--math exact]   (GPU box)
Builds a fresh Network `nets` times (new graph captures every time) and runs the 14-scene loop `reps` times on each,
comparing bit for bit with the synchronous results.  Prints one JSON line; exit code 1 on any mismatch."""
import argparse
import gc
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "3d-sis_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sis3d_synth as synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nets", type=int, default=8)
    ap.add_argument("--reps", type=int, default=25)
    ap.add_argument("--math", default="exact")
    a = ap.parse_args()
    c = synth.CASES["odd_45x27x41"]
    scenes = []
    for seed in (202, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 7, 202, 17):
        d, boxes = synth.make_scene(seed, c["dims"])
        scenes.append(synth.make_blobs(c, d, synth.make_views(seed, c["dims"], c["n_img"], boxes)))
    runs = bad = 0
    detail = []
    for n in range(a.nets):
        net, cfg = synth.make_net(c, keep_debug=False, math=a.math)
        want = []
        for b in scenes:
            P = net.forward(b, "TEST", None)
            want.append((P["rois"][0].clone(), P["cls_prob"].clone(), P["detections_host"].copy(),
                         P["mask_bits"].cpu().numpy().copy() if "mask_bits" in P else None))
        for r in range(a.reps):
            got = []
            for b, P in net.forward_pipelined(iter(scenes)):
                got.append((P["rois"][0].clone(), P["cls_prob"].clone(), P["detections_host"].copy(),
                            P["mask_bits_host"].copy() if "mask_bits_host" in P else None))
            torch.cuda.synchronize()
            runs += 1
            for k, (g, w) in enumerate(zip(got, want)):
                why = None
                if g[0].shape != w[0].shape or not torch.equal(g[0], w[0]):
                    why = "rois"
                elif not torch.equal(g[1], w[1]):
                    why = "cls_prob"
                elif not np.array_equal(g[2], w[2]):
                    why = "det"
                elif (g[3] is None) != (w[3] is None) or (g[3] is not None and not np.array_equal(g[3], w[3])):
                    why = "mask_bits"
                if why:
                    bad += 1
                    detail.append(dict(net=n, rep=r, scene=k, field=why))
        del net
        gc.collect()
        torch.cuda.synchronize()
    print(json.dumps(dict(loops=runs, scenes_per_loop=len(scenes), mismatches=bad, detail=detail[:20], math=a.math)))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
