"""Host cost of one CUDA-graph launch of the static stage (with / without parallel branches)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "3d-sis_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
torch.set_num_threads(1)
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
st = list(net._slots[0]["graphs"].values())[0]
torch.cuda.synchronize()
ts = []
for _ in range(50):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); st["graph"].replay(); ts.append((time.perf_counter() - t0) * 1e6)
torch.cuda.synchronize()
print("branches=%s kernels/graph=%d  host us per replay: median %.1f min %.1f" % (os.environ.get("SIS3D_BRANCHES", "1"), st["n_kernels"], np.median(ts), min(ts)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); st["graph"].replay(); e1.record(); torch.cuda.synchronize()
print("   gpu ms per replay %.3f" % e0.elapsed_time(e1))
