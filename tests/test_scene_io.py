"""CPU: .scene/.chunk container round trip (format of datagen/SceneSampler/main.cpp:348-395)."""
import numpy as np

from lib.datasets.scene_io import encode_tsdf, read_scene, write_scene


def test_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    sdf = rng.normal(0, 2, (12, 7, 9)).astype(np.float32)
    boxes = np.array([[1, 0, 2, 6.5, 5, 8, 4], [0, 0, 0, 3, 3, 3, 33]], np.float32)
    masks = [(4, rng.integers(0, 2, (6, 5, 6)).astype(np.uint16)), (33, rng.integers(0, 3, (3, 3, 3)).astype(np.uint16))]
    w2g = np.diag([21.3, 21.3, 21.3, 1.0]).astype(np.float32)
    w2g[:3, 3] = [3, -2, 7]
    p = tmp_path / "a.chunk"
    write_scene(p, sdf, boxes, masks, [1.0, 0.5], w2g, [20, 40, 60])
    s = read_scene(p)
    assert np.array_equal(s["sdf"], sdf) and np.array_equal(s["boxes"], boxes)
    assert all(a[0] == b[0] and np.array_equal(a[1], b[1]) for a, b in zip(s["masks"], masks))
    assert np.allclose(s["world2grid"], w2g, atol=1e-4) and list(s["frame_ids"]) == [20, 40, 60]
    assert list(s["part_in_volume"]) == [1.0, 0.5]
    # geometry-only file (no optional sections) + x-fastest storage order
    write_scene(p, sdf)
    s = read_scene(p)
    assert s["masks"] == [] and s["world2grid"] is None
    raw = np.fromfile(p, dtype="<f4", offset=24, count=3)
    assert np.array_equal(raw, sdf[:3, 0, 0])


def test_tsdf_encoding():
    sdf = np.array([[[-5.0, -1.0, -0.5, 0.2, 4.0]]], np.float32)
    enc = encode_tsdf(sdf)
    assert np.array_equal(enc[0, 0, 0], np.array([3.0, 1.0, 0.5, 0.2, 3.0], np.float32))
    assert np.array_equal(enc[1, 0, 0], [0, 0, 1, 1, 1])
