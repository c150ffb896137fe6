"""CPU: the vectorised voxel -> mesh export (3d-sis_b200/tools/scannet_benchmark/vox2mesh.py) against the
literal restatement of the reference's loops (oracle/port.py), plus the on-disk benchmark format."""
import os
import pickle

import numpy as np
import pytest

from tools.scannet_benchmark import vox2mesh as V


def _case(seed, dims=(40, 20, 36), n_box=9, n_vert=4000):
    rng = np.random.default_rng(seed)
    box, masks = [], []
    for _ in range(n_box):
        lo = np.array([rng.integers(0, d - 6) for d in dims]) + rng.choice([0.0, 0.5, 0.49, -0.3], 3)
        size = rng.integers(2, 9, 3)
        hi = np.minimum(lo + size + rng.choice([0.0, 0.5, 0.51], 3), np.array(dims) - 0.6)
        lo = np.maximum(lo, 0)
        box.append(np.concatenate([lo, hi]))
        shape = (np.rint(hi).astype(int) - np.rint(lo).astype(int)).clip(0)
        masks.append((rng.random(tuple(shape)) < 0.6).astype(np.float32))
    box = np.array(box, dtype=np.float32)
    cls = rng.integers(1, 19, n_box)
    conf = rng.uniform(0.3, 0.99, n_box).astype(np.float32)
    w2g = np.eye(4)
    w2g[:3, :3] *= 1 / 0.046875
    w2g[:3, 3] = rng.uniform(0, 3, 3)
    verts = rng.uniform(-0.2, 1.0, (n_vert, 3)) * (np.array(dims) * 0.046875) * 1.05
    return box, cls, conf, masks, w2g, verts, dims


@pytest.mark.parametrize("seed", range(5))
def test_paint_and_vertex_labels_equal_reference_loops(oracle, seed, tmp_path):
    box, cls, conf, masks, w2g, verts, dims = _case(seed)
    scene = V.paint_instances(box, cls, conf, masks, dims)
    want_scene = oracle.vox2mesh_paint(box, cls, conf, masks, dims)
    assert np.array_equal(scene, want_scene)
    assert (scene != 0).sum() > 50
    ic, im, icf = V.export(verts, w2g, scene, str(tmp_path), "scene0000_00")
    wc, wm, wcf = oracle.vox2mesh_labels(verts, w2g, want_scene)
    assert list(ic.items()) == list(wc.items())          # same instances, same (first-seen) order, same classes
    assert {k: list(v) for k, v in im.items()} == wm
    assert all(float(icf[k]) == float(wcf[k]) for k in wc)
    # files: one line per instance + a 0/1 mask per vertex
    lines = open(tmp_path / "scene0000_00.txt").read().splitlines()
    assert len(lines) == len(wc)
    for line, k in zip(lines, wc):
        f, c, s = line.split()
        assert f == f"predicted_masks/scene0000_00_{k:03d}.txt" and int(c) == wc[k] and float(s) == float(wcf[k])
        m = np.loadtxt(tmp_path / f, dtype=np.uint8)
        assert m.shape == (len(verts),) and np.array_equal(np.nonzero(m)[0], np.array(wm[k]))


def test_nn_search_scalar_matches_vectorised():
    rng = np.random.default_rng(3)
    scene = np.where(rng.random((12, 10, 11)) < 0.05, rng.integers(101, 900, (12, 10, 11)) + 0.5, 0.0)
    pts = np.array([(x, y, z) for x in range(1, 11) for y in range(1, 9) for z in range(1, 10)])
    value, valid = V.vertex_labels(pts.astype(np.float64), np.eye(4), scene)
    for (x, y, z), v, ok in zip(pts, value, valid):
        a, b, c = V.nn_search(scene, x, y, z)
        assert ok == (a != -1) and (not ok or v == scene[a, b, c])


def test_load_pred_reads_the_driver_files(tmp_path, oracle):
    box, cls, conf, masks, _, _, dims = _case(11)
    keep = np.array([True, False, True, True, False, True, True, True, False])
    np.save(tmp_path / "pred_box.npy", np.concatenate([box, np.zeros((len(box), 1), np.float32)], 1))
    np.save(tmp_path / "pred_class.npy", cls)
    np.save(tmp_path / "pred_conf.npy", conf)
    kept_masks = [m for m, k in zip(masks, keep) if k]
    pickle.dump(kept_masks, open(tmp_path / "pred_mask", "wb"))
    pickle.dump([bool(k) for k in keep], open(tmp_path / "pred_mask_index", "wb"))
    scene = V.load_pred(str(tmp_path), dims)
    assert np.array_equal(scene, oracle.vox2mesh_paint(box[keep], cls[keep], conf[keep], kept_masks, dims))


def test_load_matrix(tmp_path):
    p = tmp_path / "w2g.txt"
    p.write_text("21.3 0 0 100.5\n0 21.3 0 50\n0 0 21.3 75.25\n0 0 0 1\n")
    m = V.load_matrix(str(p))
    assert np.allclose(m[:, 3], [90.5, 34.0, 65.25, 1.0]) and np.allclose(np.diag(m)[:3], 21.3)
