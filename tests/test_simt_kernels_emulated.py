"""CPU: the SHIPPED CUDA-core kernels of csrc/conv_simt.cu (fp32 implicit-GEMM conv incl. the tap-table gather path, the tiled
MaxPool3d, weight packing, layout transposes) compiled for the host by tools/cuda_host_emu.py (one CUDA block = blockDim
threads meeting at a barrier) and driven through the same C entry points -- a second, GPU-less check of their index
arithmetic and shared-memory staging next to the GPU parity tests (tests/test_gpu_ops.py)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT
from lib import _sis3d as S

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def emu(emu_lib):
    return emu_lib


def _p(t):
    return C.c_void_p(t.data_ptr())


def _pack(emu, w):
    cout, cin, ks = w.shape[0], w.shape[1], w.shape[2]
    packed = torch.empty(ks ** 3 * cin, (cout + 3) // 4 * 4)
    assert emu.sis3d_pack_conv_weight(_p(w.contiguous()), cout, cin, ks, _p(packed), None) == 0
    return packed


@pytest.mark.parametrize("cin,cout,ks,stride,pad,dims,layout", [
    (2, 32, 2, 2, 0, (12, 10, 8), "ncdhw"),    # geometry1.0: C_in = 2, gather path with the tap table, NCDHW input
    (2, 64, 3, 1, 1, (7, 6, 9), "ncdhw"),      # mask head layer 1
    (32, 64, 2, 2, 0, (8, 6, 10), "vc"),       # fast path (C_in % 16 == 0)
    (64, 19, 1, 1, 0, (5, 7, 6), "vc"),        # mask head's 1x1 classifier (sigmoid)
    (16, 24, 3, 1, 1, (9, 5, 7), "vc")])
def test_conv3d_emulated_vs_torch(emu, cin, cout, ks, stride, pad, dims, layout):
    rng = np.random.default_rng(cin * 7 + cout)
    x = torch.from_numpy(rng.standard_normal((1, cin) + dims).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((cout, cin, ks, ks, ks)) / np.sqrt(cin * ks ** 3)).astype(np.float32))
    b = torch.from_numpy(rng.standard_normal(cout).astype(np.float32))
    act = 2 if ks == 1 else 1
    ref = F.conv3d(x, w, b, stride=stride, padding=pad)
    ref = torch.sigmoid(ref) if act == 2 else F.relu(ref)
    od = tuple(ref.shape[2:])
    nvox = int(np.prod(dims))
    if layout == "ncdhw":
        xin, in_sc, in_stride = x[0].contiguous(), nvox, [dims[1] * dims[2], dims[2], 1]
    else:
        xin, in_sc, in_stride = x[0].permute(1, 2, 3, 0).contiguous(), 1, [dims[1] * dims[2] * cin, dims[2] * cin, cin]
    reg, n_tiles = S.regions_array([0], [0], [list(dims)], [list(od)], [in_stride])
    regions = torch.from_numpy(reg.view(np.uint8).copy())
    out = torch.full(od + (cout + 4,), 7.0)
    packed = _pack(emu, w)  # keep the tensor alive across the call (ctypes only sees the address)
    rc = emu.sis3d_conv3d_ex(_p(xin), C.c_int64(in_sc), _p(packed), _p(b), None, 0, 0, _p(out), None, cout + 4, 4, _p(regions), 1,
                             n_tiles, cin, cout, ks, stride, pad, act, None)
    assert rc == 0
    got = out[..., 4:].permute(3, 0, 1, 2)
    assert torch.all(out[..., :4] == 7.0)
    torch.testing.assert_close(got, ref[0], atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("Cn,dims", [(64, (9, 5, 7)), (16, (17, 8, 3)), (8, (6, 9, 11)), (128, (8, 8, 8))])
def test_maxpool3_emulated_exact(emu, Cn, dims):
    x = torch.randn(1, Cn, *dims)
    ref = F.max_pool3d(x, 3, 1, 1)[0]
    xd = x[0].permute(1, 2, 3, 0).contiguous()
    out = torch.zeros(*dims, 2 * Cn)
    assert emu.sis3d_maxpool3(_p(xd), _p(out), 2 * Cn, Cn, *dims, Cn, None) == 0
    assert torch.equal(out[..., Cn:].permute(3, 0, 1, 2), ref) and not out[..., :Cn].any()


def test_vc_to_ncdhw_emulated(emu):
    x = torch.randn(77, 19)
    out = torch.empty(19, 77)
    assert emu.sis3d_vc_to_ncdhw(_p(x), _p(out), C.c_int64(77), 19, None) == 0
    assert torch.equal(out, x.t())
