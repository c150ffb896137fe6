"""CPU: wiring of the work-in-progress ENet executor (lib/nets/enet.py) -- strides, paddings, output sizes, residual /
pooled-skip hookup, weight packing order, channel slices -- checked against the reference features by substituting a numpy
model of the C contract declared in csrc/enet2d/sis3d_enet.h for the CUDA library.  (The CUDA kernels themselves have not
run on a GPU yet; this test pins everything above them.)"""
import ctypes as C

import numpy as np
import torch

from conftest import load_golden
from lib.nets.enet import EnetEncoder, _Conv


def _arr(ptr, n):
    return np.ctypeslib.as_array((C.c_float * int(n)).from_address(int(ptr if not hasattr(ptr, "value") else ptr.value)))


class _NumpyEnetLib:
    """The documented behaviour of libsis3d_enet.so, on host memory."""

    def __init__(self):
        self.packed_shapes = {}

    def sis3d_enet_pack_weight(self, w, cout, cin, kh, kw, packed, stream):
        ldw = (cout + 3) // 4 * 4
        src = _arr(w, cout * cin * kh * kw).reshape(cout, cin, kh * kw)
        dst = _arr(packed, kh * kw * cin * ldw).reshape(kh * kw, cin, ldw)
        dst[:] = 0
        dst[:, :, :cout] = np.transpose(src, (2, 1, 0))  # k = (ky*kw + kx)*cin + c
        return 0

    def sis3d_enet_conv2d(self, args, stream):
        a = C.cast(args, C.POINTER(_Conv)).contents
        N, H, W, cin, Ho, Wo, cout = a.N, a.H, a.W, a.cin, a.Ho, a.Wo, a.cout
        span = (N - 1) * a.in_sn + (H - 1) * a.in_sy + (W - 1) * a.in_sx + (cin - 1) * a.in_sc + 1
        flat = _arr(a.inp, span)
        x = np.lib.stride_tricks.as_strided(flat, (N, H, W, cin), tuple(4 * s for s in (a.in_sn, a.in_sy, a.in_sx, a.in_sc)))
        w = _arr(a.w, a.kh * a.kw * cin * a.ldw).reshape(a.kh, a.kw, cin, a.ldw)[..., :cout]
        acc = np.zeros((N, Ho, Wo, cout), dtype=np.float64)
        for ky in range(a.kh):
            for kx in range(a.kw):
                iy = np.arange(Ho) * a.stride - a.pad_y + ky * a.dil
                ix = np.arange(Wo) * a.stride - a.pad_x + kx * a.dil
                vy, vx = (iy >= 0) & (iy < H), (ix >= 0) & (ix < W)
                patch = np.zeros((N, Ho, Wo, cin), dtype=np.float64)
                patch[:, np.ix_(vy, vx)[0], np.ix_(vy, vx)[1]] = x[:, iy[vy]][:, :, ix[vx]]
                acc += patch @ w[ky, kx].astype(np.float64)
        if a.bias:
            acc += _arr(a.bias, cout)
        if a.res:
            if a.res_pool:
                r = _arr(a.res, N * 2 * Ho * 2 * Wo * a.res_ld).reshape(N, Ho, 2, Wo, 2, a.res_ld)[..., :a.res_c]
                acc[..., :a.res_c] += r.max(axis=(2, 4))
            else:
                acc[..., :a.res_c] += _arr(a.res, N * Ho * Wo * a.res_ld).reshape(N, Ho, Wo, a.res_ld)[..., :a.res_c]
        if a.slope:
            sl = _arr(a.slope, cout)
            acc = np.where(acc >= 0, acc, acc * sl)
        out = _arr(a.out, N * Ho * Wo * a.out_ld).reshape(N, Ho, Wo, a.out_ld)
        out[..., a.out_coff:a.out_coff + cout] = acc.astype(np.float32)
        return 0

    def sis3d_enet_pool_affine(self, inp, sn, sy, sx, sc, N, H, W, Cn, scale, shift, slope, out, out_ld, out_coff, stream):
        span = (N - 1) * sn + (H - 1) * sy + (W - 1) * sx + (Cn - 1) * sc + 1
        x = np.lib.stride_tricks.as_strided(_arr(inp, span), (N, H, W, Cn), tuple(4 * s for s in (sn, sy, sx, sc)))
        p = x[:, :H // 2 * 2, :W // 2 * 2].reshape(N, H // 2, 2, W // 2, 2, Cn).max(axis=(2, 4))
        v = p * _arr(scale, Cn) + _arr(shift, Cn)
        v = np.where(v >= 0, v, v * _arr(slope, Cn))
        _arr(out, N * (H // 2) * (W // 2) * out_ld).reshape(N, H // 2, W // 2, out_ld)[..., out_coff:out_coff + Cn] = v
        return 0

    def sis3d_enet_to_nchw(self, inp, ld, coff, N, P, Cn, out, stream):
        P = int(P.value if hasattr(P, "value") else P)
        x = _arr(inp, N * P * ld).reshape(N, P, ld)[..., coff:coff + Cn]
        _arr(out, N * Cn * P).reshape(N, Cn, P)[:] = np.transpose(x, (0, 2, 1))
        return 0


def test_executor_wiring_reproduces_reference_features():
    g = load_golden("enet_encoder.npz")
    params = [torch.from_numpy(g[k]) for k in sorted(k for k in g if k.startswith("p"))]
    enc = EnetEncoder(params, "cpu", lib=_NumpyEnetLib())
    x = torch.from_numpy(np.random.default_rng(int(g["seed"])).standard_normal((1, 3, 256, 328)).astype(np.float32))
    y = enc(x)
    ref = torch.from_numpy(g["features"])
    assert y.shape == ref.shape
    assert float((y - ref).abs().max()) < 2e-4 and float((y - ref).norm() / ref.norm()) < 2e-6


def test_cuda_source_under_host_emulation_matches_the_contract_model():
    """The SAME enet2d.cu that nvcc compiles for sm_100a, compiled for the host (csrc/enet2d/host_emu.h: one CUDA block =
    blockDim std::threads meeting at a barrier for __syncthreads) and driven by the executor: checks the kernels' index
    arithmetic, tap tables, shared-memory staging, epilogues and packing on a small image against the numpy contract model
    and the torch program -- everything short of running on the GPU itself."""
    import os
    import subprocess
    from conftest import ROOT
    from lib.nets import enet_program as E
    d = os.path.join(ROOT, "3d-sis_b200", "csrc", "enet2d")
    subprocess.check_call(["make", "-C", d, "emu"], stdout=subprocess.DEVNULL)
    emu = C.CDLL(os.path.join(ROOT, "3d-sis_b200", "lib", "libsis3d_enet_emu.so"))
    g = load_golden("enet_encoder.npz")
    params = [torch.from_numpy(g[k]) for k in sorted(k for k in g if k.startswith("p"))]
    x = torch.from_numpy(np.random.default_rng(5).standard_normal((2, 3, 48, 72)).astype(np.float32))
    y_emu = EnetEncoder(params, "cpu", lib=emu)(x)
    y_model = EnetEncoder(params, "cpu", lib=_NumpyEnetLib())(x)
    with torch.no_grad():
        y_ref = E.run_program(E.compile_enet(params), x)
    assert y_emu.shape == y_ref.shape == (2, 128, 6, 9)
    assert float((y_emu - y_model).abs().max()) < 1e-4
    assert float((y_emu - y_ref).abs().max()) < 2e-4 and float((y_emu - y_ref).norm() / y_ref.norm()) < 1e-5
