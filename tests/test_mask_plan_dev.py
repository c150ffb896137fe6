"""CPU: the work-in-progress DEVICE planner of the mask stage (csrc/wip/mask_plan_dev.cu, round 2: lets the mask stage join
the CUDA graph) run under host emulation and compared table by table with the native host planner sis3d_mask_plan_build."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from lib import _sis3d as S
from test_mask_plan import _dets, _plan


class PlanDev(C.Structure):
    _fields_ = [("n_kept", C.c_int32), ("canvas", C.c_int32 * 3), ("n_tiles_tc", C.c_int32), ("tiles_first", C.c_int32),
                ("overflow", C.c_int32), ("reserved", C.c_int32), ("total_voxels", C.c_int64)]


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "3d-sis_b200", "csrc", "wip"), "emu"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(ROOT, "3d-sis_b200", "lib", "libsis3d_wip_emu.so"))
    lib.sis3d_mask_plan_device_bytes.restype = C.c_size_t
    return lib


def _run(emu, det, dims, ncls, kcap, tcap, cy=0, cz=0):
    table = np.zeros((max(len(det), 1), 16), dtype=np.float32)
    table[:len(det)] = det
    table[0, 15] = len(det)  # the RoI count travels in row 0 / column 15
    blob = np.full(emu.sis3d_mask_plan_device_bytes(kcap, tcap), 0xAB, dtype=np.uint8)
    plan = PlanDev()
    rc = emu.sis3d_mask_plan_device(C.c_void_p(table.ctypes.data), len(table), *dims, ncls, cy, cz, kcap, tcap,
                                    C.c_void_p(blob.ctypes.data), C.byref(plan), None)
    assert rc == 0
    offs = (C.c_int64 * 7)()
    emu.sis3d_mask_plan_device_layout(kcap, tcap, offs)
    return plan, blob, list(offs)


@pytest.mark.parametrize("seed", range(8))
def test_device_planner_equals_host_planner(emu, seed):
    rng = np.random.default_rng(100 + seed)
    dims, ncls = (45, 27, 41), 19
    det = _dets(rng, int(rng.integers(1, 60)), dims)
    rc, hp, hblob = _plan(det, dims, ncls, 1)
    assert rc == 0
    dp, dblob, (o_first, o_last, o_tiles, o_offs, o_cls, o_kept, o_sizes) = _run(emu, det, dims, ncls, 64, 8192)
    nk = hp.n_kept
    assert (dp.n_kept, tuple(dp.canvas), dp.n_tiles_tc, dp.tiles_first, dp.total_voxels, dp.overflow) == \
        (nk, tuple(hp.canvas) if nk else (0, 0, 0), hp.n_tiles_tc, hp.tiles_first, hp.total_voxels, 0)
    if nk == 0:
        return
    rb = S.REGION_BYTES * nk
    assert np.array_equal(dblob[o_first:o_first + rb], hblob[hp.off_first:hp.off_first + rb])
    assert np.array_equal(dblob[o_last:o_last + rb], hblob[hp.off_last:hp.off_last + rb])
    assert np.array_equal(dblob[o_tiles:o_tiles + 32 * hp.n_tiles_tc], hblob[hp.off_rest:hp.off_rest + 32 * hp.n_tiles_tc])
    assert np.array_equal(dblob[o_offs:o_offs + 8 * (nk + 1)], hblob[hp.off_offs:hp.off_offs + 8 * (nk + 1)])
    for o_dev, o_host, n in ((o_cls, hp.off_cls, 4 * nk), (o_kept, hp.off_kept, 4 * nk), (o_sizes, hp.off_sizes, 12 * nk)):
        assert np.array_equal(dblob[o_dev:o_dev + n], hblob[o_host:o_host + n])


def test_device_planner_fixed_canvas_and_overflow(emu):
    rng = np.random.default_rng(7)
    dims = (45, 27, 41)
    det = _dets(rng, 30, dims)
    det[:, 8] = 1.0
    p, blob, offs = _run(emu, det, dims, 19, 64, 8192, cy=27, cz=41)   # static canvas extents for graph capture
    assert tuple(p.canvas)[1:] == (27, 41) and p.overflow == 0
    first = blob[offs[0]:offs[0] + S.REGION_BYTES * p.n_kept].view(S.REGION_DTYPE)
    assert np.all(first["out_stride"] == np.array([27 * 41 * 64, 41 * 64, 64]))
    p2, _, _ = _run(emu, det, dims, 19, 8, 8192)                        # more kept RoIs than capacity
    assert p2.overflow == 1 and p2.n_kept == 8 and p2.n_tiles_tc == 0
    p3, _, _ = _run(emu, det, dims, 19, 64, 4)                          # more bricks than capacity
    assert p3.overflow == 1 and p3.n_tiles_tc == 0


@pytest.mark.parametrize("seed", range(4))
def test_shell_zero_makes_every_out_of_crop_neighbour_zero(emu, seed):
    """Static-canvas companion kernel (csrc/wip/mask_shell_zero.cu) under host emulation: on a garbage-filled canvas every
    voxel that a 3x3x3 tap of a crop voxel can touch outside its own crop reads zero afterwards; crop interiors are untouched."""
    rng = np.random.default_rng(40 + seed)
    nk = int(rng.integers(1, 9))
    sizes = rng.integers(1, 9, (nk, 3)).astype(np.int32)
    Yc, Zc = int(sizes[:, 1].max() + rng.integers(0, 3)), int(sizes[:, 2].max() + rng.integers(0, 3))
    Xc = int((sizes[:, 0] + 1).sum())
    row = 32  # bytes per voxel row in this test
    canvas = np.full((Xc, Yc, Zc, row), 0xFF, dtype=np.uint8)
    n_kept = np.array([nk], dtype=np.int32)
    rc = emu.sis3d_mask_shell_zero(C.c_void_p(n_kept.ctypes.data), C.c_void_p(sizes.ctypes.data), 16, Xc, Yc, Zc, row,
                                   C.c_void_p(canvas.ctypes.data), None)
    assert rc == 0
    zero = ~canvas.any(axis=3)
    xoff = np.concatenate([[0], np.cumsum(sizes[:, 0] + 1)[:-1]])
    inside = np.zeros((Xc, Yc, Zc), dtype=bool)
    for j in range(nk):
        w, h, l = sizes[j]
        inside[xoff[j]:xoff[j] + w, :h, :l] = True
    assert not (zero & inside).any(), "crop interiors must not be touched"
    for j in range(nk):
        w, h, l = (int(v) for v in sizes[j])
        x0 = int(xoff[j])
        lo = np.array([max(x0 - 1, 0), 0, 0])
        hi = np.array([min(x0 + w + 1, Xc), min(h + 1, Yc), min(l + 1, Zc)])
        halo = np.zeros((Xc, Yc, Zc), dtype=bool)
        halo[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = True   # everything a tap of this crop can read inside the canvas
        halo[x0:x0 + w, :h, :l] = False                       # minus the crop itself
        assert zero[halo].all(), f"crop {j}: a readable out-of-crop voxel is not zero"
