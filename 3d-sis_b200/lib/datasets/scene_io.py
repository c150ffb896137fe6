"""Binary `.scene` / `.chunk` container of 3D-SIS (written by datagen/SceneSampler/main.cpp:348-395,
parsed by lib/datasets/dataset.py:45-187 in the reference).  numpy-based reader + a writer (used to
produce synthetic files for tests and benchmarks; the datasets themselves are not redistributable).

Layout (little endian):
  u64 X, Y, Z | f32 sdf[X*Y*Z] (x fastest) | u32 n_box | n_box x (f32 min xyz, f32 max xyz, u32 label)
  | u32 n_mask | n_mask x (u32 label, u64 X,Y,Z, u16 data[X*Y*Z] (x fastest))
  | u32 n_box | n_box x f32 part_in_volume
  | f32 grid2world[16] (row-major; the reference inverts it to world2grid) | u32 n_img | n_img x u32 frame id
The last three sections are optional (present when the sampler was run with images / keep-threshold).
"""
from __future__ import annotations

import numpy as np


class _Cursor:
    def __init__(self, buf):
        self.buf, self.pos = buf, 0

    def take(self, dtype, count=1):
        dt = np.dtype(dtype)
        n = dt.itemsize * count
        if self.pos + n > len(self.buf):
            raise EOFError("not enough bytes in file to satisfy read request")
        out = np.frombuffer(self.buf, dtype=dt, count=count, offset=self.pos)
        self.pos += n
        return out

    @property
    def eof(self):
        return self.pos >= len(self.buf)


def read_scene(path, raw_sdf=False):
    """-> dict(sdf[X,Y,Z] f32, boxes[n,7] (min xyz, max xyz, raw label), masks [list of (label, u16[X,Y,Z])],
    part_in_volume[n] | None, world2grid[4,4] | None, frame_ids[n_img] | None).
    raw_sdf=True: `sdf` is left as stored in the file -- a flat float32 array, x fastest -- and `dims` = (X, Y, Z) is added: the
    layout sis3d_chunk_decode consumes on the device (no host-side transposition or TSDF encoding)."""
    with open(path, "rb") as f:
        cur = _Cursor(f.read())
    X, Y, Z = (int(v) for v in cur.take("<u8", 3))
    flat = cur.take("<f4", X * Y * Z)
    sdf = flat if raw_sdf else flat.reshape((X, Y, Z), order="F").astype(np.float32)
    n_box = int(cur.take("<u4")[0])
    boxes = np.zeros((n_box, 7), dtype=np.float32)
    for i in range(n_box):
        boxes[i, :6] = cur.take("<f4", 6)
        boxes[i, 6] = cur.take("<u4")[0]
    out = dict(sdf=sdf, dims=(X, Y, Z), boxes=boxes, masks=[], part_in_volume=None, world2grid=None, frame_ids=None)
    if cur.eof:
        return out
    n_mask = int(cur.take("<u4")[0])
    for _ in range(n_mask):
        label = int(cur.take("<u4")[0])
        mx, my, mz = (int(v) for v in cur.take("<u8", 3))
        out["masks"].append((label, cur.take("<u2", mx * my * mz).reshape((mx, my, mz), order="F").copy()))
    if cur.eof:
        return out
    nb = int(cur.take("<u4")[0])
    out["part_in_volume"] = cur.take("<f4", nb).copy()
    if cur.eof:
        return out
    g2w = cur.take("<f4", 16).reshape(4, 4).astype(np.float32)
    out["world2grid"] = np.linalg.inv(g2w).astype(np.float32)  # dataset.py:140
    n_img = int(cur.take("<u4")[0])
    out["frame_ids"] = cur.take("<u4", n_img).copy() if n_img and not cur.eof else np.zeros(0, np.uint32)
    return out


def write_scene(path, sdf, boxes=None, masks=(), part_in_volume=None, world2grid=None, frame_ids=None):
    sdf = np.asarray(sdf, dtype=np.float32)
    boxes = np.zeros((0, 7), np.float32) if boxes is None else np.asarray(boxes, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(np.asarray(sdf.shape, dtype="<u8").tobytes())
        f.write(sdf.reshape(-1, order="F").astype("<f4").tobytes())
        f.write(np.uint32(len(boxes)).tobytes())
        for b in boxes:
            f.write(b[:6].astype("<f4").tobytes())
            f.write(np.uint32(int(b[6])).tobytes())
        if not masks and part_in_volume is None and world2grid is None:
            return
        f.write(np.uint32(len(masks)).tobytes())
        for label, m in masks:
            m = np.asarray(m)
            f.write(np.uint32(label).tobytes())
            f.write(np.asarray(m.shape, dtype="<u8").tobytes())
            f.write(m.reshape(-1, order="F").astype("<u2").tobytes())
        if part_in_volume is None and world2grid is None:
            return
        piv = np.ones(len(boxes), np.float32) if part_in_volume is None else np.asarray(part_in_volume, np.float32)
        f.write(np.uint32(len(piv)).tobytes())
        f.write(piv.astype("<f4").tobytes())
        if world2grid is None:
            return
        f.write(np.linalg.inv(np.asarray(world2grid, np.float64)).astype("<f4").tobytes())
        ids = np.zeros(0, np.uint32) if frame_ids is None else np.asarray(frame_ids, np.uint32)
        f.write(np.uint32(len(ids)).tobytes())
        f.write(ids.astype("<u4").tobytes())


def encode_tsdf(sdf, truncation=3.0):
    """2-channel network input [2,X,Y,Z]: |clip(sdf)|, sdf > -1  (dataset.py:66-68)."""
    t = np.abs(np.clip(sdf, -truncation, truncation))
    return np.stack([t, (sdf > -1).astype(np.float32)], 0).astype(np.float32)
