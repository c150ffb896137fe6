"""Round-2 starting point: run the WIP ENet executor (lib/nets/enet.py over libsis3d_enet.so) on a B200 and
compare with the features of the unmodified reference ENet (tests/golden/enet_encoder.npz).  Not part of the test suite:
the kernel has not been validated yet.    gpurun -- 'python tools/enet_check.py'"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-sis_b200"))
import numpy as np
import torch

from lib.nets.enet import EnetEncoder

g = dict(np.load(os.path.join(ROOT, "tests", "golden", "enet_encoder.npz")))
params = [torch.from_numpy(g[k]) for k in sorted(k for k in g if k.startswith("p"))]
x = torch.from_numpy(np.random.default_rng(int(g["seed"])).standard_normal((1, 3, 256, 328)).astype(np.float32)).cuda()
enc = EnetEncoder(params, "cuda:0")
y = enc(x)
torch.cuda.synchronize()
ref = torch.from_numpy(g["features"]).cuda()
print("shape", tuple(y.shape), "max abs err", float((y - ref).abs().max()), "rel", float((y - ref).norm() / ref.norm()))
xb = x.repeat(5, 1, 1, 1)
for _ in range(3):
    enc(xb)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    enc(xb)
torch.cuda.synchronize()
print("5 images: %.3f ms per call (eager launches, no graph)" % ((time.perf_counter() - t0) / 20 * 1e3))
