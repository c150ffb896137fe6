"""Kernel timeline of one forward (CUDA-graph replay + ragged mask stage) via torch.profiler/CUPTI."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "3d-sis_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
torch.set_num_threads(1)
import sis3d_synth as _synth
from sis3d_synth import make_net
from sis3d_synth import CASES
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda", 0)
net, cfg = make_net(CASES["cfg2_96x48x96"], keep_debug=False, math=os.environ.get("SIS3D_CONV_MATH", "exact"))
data, _boxes = _synth.make_scene(1000, (96, 48, 96))
views = _synth.make_views(1000, (96, 48, 96), 5, _boxes)
blobs = {"data": torch.from_numpy(data).to(dev), "id": ["x"],
         "nearest_images": {"images": [torch.from_numpy(views["feats"]).to(dev)], "depths": [torch.from_numpy(views["depths"]).to(dev)],
                            "poses": [torch.from_numpy(views["poses"])], "world2grid": [torch.from_numpy(views["world2grid"])]}}
for _ in range(4):
    net.forward(blobs, "TEST", None)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        net.forward(blobs, "TEST", None)
    torch.cuda.synchronize()
out = os.path.join(ROOT, "gpurun_out", "trace_replay.json")
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
# last forward = last third of the events
n = len(ev) // 3
last = ev[-n:]
t0 = last[0]["ts"]
busy = sum(e["dur"] for e in last)
span = last[-1]["ts"] + last[-1]["dur"] - t0
print("events per forward: %d  span %.1f us  busy %.1f us  (gaps %.1f us)" % (n, span, busy, span - busy))
prev_end = t0
rows = []
for e in last:
    rows.append((e["ts"] - t0, e["dur"], e["ts"] - prev_end, e["name"].split("(")[0].replace("void ", "")[:60]))
    prev_end = e["ts"] + e["dur"]
with open(os.path.join(ROOT, "gpurun_out", "timeline_one_forward.txt"), "w") as f:
    f.write("# start_us  dur_us  gap_before_us  kernel   (one forward: graph replay + mask stage; B200, tf32)\n")
    f.write("# events %d span %.1f us busy %.1f us\n" % (n, span, busy))
    for r in rows:
        f.write("%9.1f %8.1f %8.1f  %s\n" % r)
for r in rows:
    print("%9.1f %8.1f %8.1f  %s" % r)
