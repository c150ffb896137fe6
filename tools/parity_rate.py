#!/usr/bin/env python
"""Per math mode: fraction of scenes whose integer outputs (proposal count/order, level ids, class argmax, mask-keep flags,
crop bounds) equal the fp32 CUDA-core path exactly -- over the 24 bench chunks (seeds 1000..1023, cfg2) and the golden cases.
Run on a GPU box:  python tools/parity_rate.py [--chunks 24] [--modes exact mixed tf32 fp16 tf32x3] -> JSON on stdout."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "3d-sis_b200"), ROOT):
    sys.path.insert(0, p)

import sis3d_synth as synth  # noqa: E402
from lib.utils.parity import parity_rate  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=24)
    ap.add_argument("--seed0", type=int, default=1000)
    ap.add_argument("--modes", nargs="+", default=["exact", "mixed", "tf32", "fp16", "tf32x3"])
    args = ap.parse_args()
    out = {}
    # (a) bench chunks: cfg2 shape, the seeds bench.py rotates on rank 0
    c = synth.CASES["cfg2_96x48x96"]
    blobs = []
    for j in range(args.chunks):
        d, boxes = synth.make_scene(args.seed0 + j, c["dims"])
        blobs.append(synth.make_blobs(c, d, synth.make_views(args.seed0 + j, c["dims"], c["n_img"], boxes)))
    mk = lambda mode: synth.make_net(c, keep_debug=False, math=mode)[0]  # noqa: E731
    ref = None
    out["bench_chunks"] = {}
    for m in args.modes:
        r, ref = parity_rate(mk, blobs, m, ref_sigs=ref)
        out["bench_chunks"][m] = r
        print(json.dumps({"set": "bench_chunks", **r}), file=sys.stderr, flush=True)
    # (b) the golden cases (different shapes / configs)
    out["golden_cases"] = {m: dict(scenes=0, exact_scenes=0, fields={}) for m in args.modes}
    for tag, cc in synth.CASES.items():
        d, boxes = synth.make_scene(cc["seed"], cc["dims"])
        v = synth.make_views(cc["seed"], cc["dims"], cc["n_img"], boxes) if cc["use_images"] else None
        bl = [synth.make_blobs(cc, d, v)]
        mk2 = lambda mode, cc=cc: synth.make_net(cc, keep_debug=False, math=mode)[0]  # noqa: E731
        ref2 = None
        for m in args.modes:
            r, ref2 = parity_rate(mk2, bl, m, ref_sigs=ref2)
            g = out["golden_cases"][m]
            g["scenes"] += 1
            g["exact_scenes"] += r["exact_scenes"]
            for k, n in r["first_mismatch_fields"].items():
                g["fields"][f"{tag}:{k}"] = n
    print(json.dumps(out))


if __name__ == "__main__":
    main()
