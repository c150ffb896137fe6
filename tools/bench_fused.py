"""Device time of the fused bottleneck tail vs the two-kernel path (CUDA events, back-to-back launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-sis_b200"))
import torch
from lib import _sis3d as S
DEV = torch.device("cuda", 0)
for cin, cmid, cout, dims in [(32, 32, 64, (48, 24, 48)), (32, 32, 32, (48, 24, 48)), (64, 64, 128, (24, 12, 24)), (32, 32, 64, (24, 12, 24))]:
    x = torch.randn(*dims, cin, device=DEV); r = torch.randn(*dims, cout, device=DEV)
    w2 = torch.randn(cmid, 27 * cin, device=DEV) * 0.05; w3 = torch.randn(cout, cmid, device=DEV) * 0.1
    mid = torch.empty(*dims, cmid, device=DEV); out = torch.empty(*dims, cout, device=DEV)
    st = S.stream()
    def fused():
        S.check(S.lib.sis3d_conv3d_k3_tc_fused(S.ptr(x), S.ptr(w2), None, S.ptr(w3), None, S.ptr(r), cout, 0, S.ptr(out), cout, 0, *dims, cin, cmid, cout, 1, st))
    def two():
        S.check(S.lib.sis3d_conv3d_k3_tc(S.ptr(x), S.ptr(w2), None, None, 0, 0, S.ptr(mid), cmid, 0, *dims, cin, cmid, 3, None, 0, 1, st))
        S.check(S.lib.sis3d_conv3d_k3_tc(S.ptr(mid), S.ptr(w3), None, S.ptr(r), cout, 0, S.ptr(out), cout, 0, *dims, cmid, cout, 1, None, 0, 1, st))
    for name, f in (("fused", fused), ("two", two)):
        for _ in range(20): f()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200): f()
        b.record(); torch.cuda.synchronize()
        print(cin, cmid, cout, dims, name, "%.2f us" % (a.elapsed_time(b) * 5))
