#!/usr/bin/env python
"""DRAM traffic of the benchmarked build from an ncu launch list, written where bench.py looks for it.

  (GPU box)  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 4000 --csv \\
                 --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 8 --lean
  (here)     python tools/ncu_traffic.py gpurun_out/traffic.csv   ->  profiles/r2_traffic.json

Per-scene traffic = dram bytes of every libsis3d launch between consecutive `detect_decode_kernel` launches (one per scene),
averaged over the steady-state scenes of the second half of the capture (caches are flushed per kernel under ncu, so this is
an upper bound of the live traffic).  The file carries the hash of the kernel sources; bench.py reports `traffic` only when
that hash equals the build it is running."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    path = sys.argv[1]
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    launches = {}
    for r in csv.DictReader(lines):
        k = int(r["ID"])
        d = launches.setdefault(k, {"name": re.sub(r"^void ", "", r["Kernel Name"]), "bytes": 0.0, "us": 0.0})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        if r["Metric Name"].startswith("dram__bytes"):
            d["bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        elif r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if unit in ("ns", "nsecond") else v
    seq = [launches[k] for k in sorted(launches)]
    marks = [i for i, d in enumerate(seq) if "detect_decode_kernel" in d["name"]]
    half = marks[len(marks) // 2:]
    if len(half) < 3:
        raise SystemExit("too few scenes in the capture")
    per_scene = [sum(d["bytes"] for d in seq[a + 1:b + 1] if "sis3d::" in d["name"]) for a, b in zip(half[:-1], half[1:])]
    us_scene = [sum(d["us"] for d in seq[a + 1:b + 1] if "sis3d::" in d["name"]) for a, b in zip(half[:-1], half[1:])]
    n_launch = [sum(1 for d in seq[a + 1:b + 1] if "sis3d::" in d["name"]) for a, b in zip(half[:-1], half[1:])]
    rpn = [d["bytes"] for d in seq[half[0]:] if re.search(r"conv3d_k3_tc_kernel<128, 3, 4, \d+, 2, 0, [12](, \d+)?>", d["name"])]
    import bench
    out = {"sources_sha": bench.sources_sha(), "capture": os.path.basename(path), "scenes_averaged": len(per_scene),
           "forward_dram_bytes_per_scene": sum(per_scene) / len(per_scene),
           "serialised_kernel_us_per_scene": sum(us_scene) / len(us_scene), "launches_per_scene": sum(n_launch) / len(n_launch),
           "rpn_kernel_dram_bytes_per_launch": (sum(rpn) / len(rpn)) if rpn else None,
           "note": "ncu replays each kernel with cold caches: an upper bound of the live DRAM traffic"}
    dst = os.path.join(ROOT, "profiles", "r2_traffic.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
