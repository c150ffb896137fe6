"""Diagnostic: where does a forward spend its time?  (run on the GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "3d-sis_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
if os.environ.get("SIS3D_HOST_THREADS"):
    torch.set_num_threads(int(os.environ["SIS3D_HOST_THREADS"]))
print("torch threads", torch.get_num_threads(), "cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
import sis3d_synth as _synth
from sis3d_synth import make_net
from sis3d_synth import CASES

dev = torch.device("cuda", 0)
net, cfg = make_net(CASES["cfg2_96x48x96"], keep_debug=False, math=os.environ.get("SIS3D_CONV_MATH", "exact"))
data, _boxes = _synth.make_scene(1000, (96, 48, 96))
views = _synth.make_views(1000, (96, 48, 96), 5, _boxes)
blobs = {"data": torch.from_numpy(data).to(dev), "id": ["x"],
         "nearest_images": {"images": [torch.from_numpy(views["feats"]).to(dev)], "depths": [torch.from_numpy(views["depths"]).to(dev)],
                            "poses": [torch.from_numpy(views["poses"])], "world2grid": [torch.from_numpy(views["world2grid"])]}}
for _ in range(3):
    net.forward(blobs, "TEST", None)
torch.cuda.synchronize()

def wall(fn, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

st = list(net._slots[0]["graphs"].values())[0]
print("graph replay + sync        : %.3f ms" % wall(lambda: (st["graph"].replay(), torch.cuda.synchronize())))
det = net._predictions["detections_host"]; n = det.shape[0]
print("mask branch + sync (n=%d)  : %.3f ms" % (n, wall(lambda: (net._mask_branch(st["scene"], det, n), torch.cuda.synchronize()))))
print("full forward               : %.3f ms" % wall(lambda: net.forward(blobs, "TEST", None)))
net._use_graph = False
print("full forward (eager)       : %.3f ms" % wall(lambda: net.forward(blobs, "TEST", None)))
# eager static stage with per-op events (GPU-side durations incl. launch gaps)
net._prof = {}
net.forward(blobs, "TEST", None); torch.cuda.synchronize()
tot = 0.0
for k, v in sorted(net._prof.items(), key=lambda kv: -sum(a.elapsed_time(b) for a, b in kv[1]))[:12]:
    ms = sum(a.elapsed_time(b) for a, b in v); tot += ms
    print("   %-40s %.3f ms" % (k, ms))
print("   sum of all ops %.3f ms" % sum(sum(a.elapsed_time(b) for a, b in v) for v in net._prof.values()))
net._prof = None
# graph replay timed by events only
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); st["graph"].replay(); e1.record(); torch.cuda.synchronize()
print("graph replay (events)      : %.3f ms" % e0.elapsed_time(e1))

# ---- step-by-step replica of Network.forward (graph mode) with a sync after each step
net._use_graph = True
from lib.layer_utils import projection as proj
from lib.utils.config import cfg
imgs = blobs["nearest_images"]
def T(label, fn, n=10):
    ts = []
    out = None
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("   %-34s %.3f ms (min %.3f)" % (label, float(np.mean(ts)), min(ts)))
    return out
T("_ensure_packed", net._ensure_packed)
vp = T("view_params", lambda: proj.view_params(cfg.INTRINSIC, (41, 32), cfg.PROJ_DEPTH_MIN, cfg.PROJ_DEPTH_MAX, (96, 48, 96), None, imgs["poses"][0], imgs["world2grid"][0]))
T("copy scene", lambda: st["scene"].copy_(blobs["data"], non_blocking=True))
T("copy feats", lambda: st["feats"].copy_(imgs["images"][0], non_blocking=True))
T("copy depths", lambda: st["depths"].copy_(imgs["depths"][0], non_blocking=True))
T("copy vp (pageable H2D)", lambda: st["vp"].copy_(vp, non_blocking=True))
T("replay", lambda: st["graph"].replay())
outs = T("clones", lambda: {k: v.clone() for k, v in st["outs"].items()})
n = T("num.item()", lambda: int(outs["num"].item()))
det_host = T("det.cpu()", lambda: outs["det"][:n].cpu().numpy())
T("mask branch", lambda: net._mask_branch(st["scene"], det_host, n))
def seq():
    st["scene"].copy_(blobs["data"], non_blocking=True); st["feats"].copy_(imgs["images"][0], non_blocking=True)
    st["depths"].copy_(imgs["depths"][0], non_blocking=True); st["vp"].copy_(vp, non_blocking=True)
    st["graph"].replay()
    o = {k: v.clone() for k, v in st["outs"].items()}
    nn = int(o["num"].item())
    dh = o["det"][:nn].cpu().numpy()
    return net._mask_branch(st["scene"], dh, nn)
T("all of the above, one sync", seq)
T("net.forward", lambda: net.forward(blobs, "TEST", None))

ts = []
for _ in range(60):
    torch.cuda.synchronize(); t0 = time.perf_counter(); net.forward(blobs, "TEST", None); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
ts = np.array(ts)
print("net.forward x60: mean %.3f median %.3f min %.3f max %.3f  | >4ms: %d" % (ts.mean(), np.median(ts), ts.min(), ts.max(), (ts > 4).sum()))
print("   trace:", " ".join("%.1f" % t for t in ts[:30]))
