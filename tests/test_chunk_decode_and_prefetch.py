"""Row f3, the part in front of the path: (CPU) the device chunk-decode kernel (csrc/io.cu) under host emulation against the host
encoder, which tests/test_dataset_golden.py pins to the reference reader; the thread prefetcher; (GPU) the same kernel on the
device, and the pipelined inference driver writing exactly the files the synchronous one writes."""
import ctypes as C
import os
import pickle

import numpy as np
import pytest
import torch

from lib.datasets.prefetch import Prefetcher
from lib.datasets.scene_io import encode_tsdf, read_scene, write_scene


def _chunk(tmp_path, dims=(37, 53, 29), seed=3):
    rng = np.random.default_rng(seed)
    sdf = rng.normal(0, 2.5, dims).astype(np.float32)
    sdf.reshape(-1)[::7] = -1.0      # exactly on the `> -1` threshold
    sdf.reshape(-1)[::11] = 3.0      # exactly on the clip bound
    p = str(tmp_path / "s__0.chunk")
    write_scene(p, sdf, world2grid=np.eye(4), frame_ids=[0])
    return p, sdf


def test_chunk_decode_emulated_equals_host_encoder(emu_lib, tmp_path):
    p, sdf = _chunk(tmp_path)
    s = read_scene(p, raw_sdf=True)
    X, Y, Z = s["dims"]
    assert (X, Y, Z) == sdf.shape and s["sdf"].ndim == 1
    raw = torch.from_numpy(np.array(s["sdf"], dtype=np.float32))
    for y_keep in (48, 480):
        Yk = min(Y, y_keep)
        out = torch.full((2, X, Yk, Z), 7.0)
        rc = emu_lib.sis3d_chunk_decode(C.c_void_p(raw.data_ptr()), X, Y, Z, y_keep, C.c_float(3.0), C.c_void_p(out.data_ptr()), None)
        assert rc == 0
        want = encode_tsdf(read_scene(p)["sdf"], 3.0)[:, :, :y_keep, :]
        assert np.array_equal(out.numpy(), want)
    assert emu_lib.sis3d_chunk_decode(None, X, Y, Z, 48, C.c_float(3.0), C.c_void_p(out.data_ptr()), None) != 0


def test_prefetcher_keeps_order_and_propagates_errors():
    assert list(Prefetcher(range(100), depth=3)) == list(range(100))
    assert list(Prefetcher([], depth=2)) == []

    def bad():
        yield 1
        yield 2
        raise ValueError("reader failed")
    got = []
    with pytest.raises(ValueError):
        for v in Prefetcher(bad(), depth=2):
            got.append(v)
    assert got == [1, 2]


@pytest.mark.gpu
def test_device_decode_and_pipelined_driver_equal_the_synchronous_path(oracle, tmp_path):
    import sis3d_synth as synth
    from lib.datasets.dataset import Dataset, collate_fn
    from lib.model.trainval import run_scenes, scene_key
    c = synth.CASES["odd_45x27x41"]
    net, cfg = synth.make_net(c, keep_debug=False, math="exact")
    paths, provs = [], {}
    for k, seed in enumerate((202, 7, 8, 9, 10, 11, 12)):
        data, boxes = synth.make_scene(seed, c["dims"])
        views = synth.make_views(seed, c["dims"], c["n_img"], boxes)
        sdf = np.where(data[0, 1] > 0.5, data[0, 0], -data[0, 0]).astype(np.float32)
        sdf[(data[0, 1] < 0.5) & (data[0, 0] < 1.0)] = -1.0
        p = str(tmp_path / f"scene{k:04d}_00__0.chunk")
        write_scene(p, sdf, world2grid=views["world2grid"], frame_ids=[0, 20, 40])
        paths.append(p)
        provs[p] = views
    provider = lambda p, ids, w2g, dims: {"images": provs[p]["feats"], "depths": provs[p]["depths"], "poses": provs[p]["poses"],
                                          "world2grid": provs[p]["world2grid"]}
    host = Dataset(paths, "test", view_provider=provider)
    devd = Dataset(paths, "test", view_provider=provider, device_decode=True)
    for i in range(len(paths)):  # device decode == host encode, exactly
        a, b = collate_fn([host[i]]), collate_fn([devd[i]])
        assert b["data"].is_cuda and torch.equal(b["data"].cpu(), a["data"])
    out_a, out_b = str(tmp_path / "sync"), str(tmp_path / "pipe")
    assert run_scenes(net, (collate_fn([host[i]]) for i in range(len(paths))), out_a, pipelined=False) == len(paths)
    from lib.datasets.prefetch import Prefetcher as PF
    loader = (collate_fn([item]) for item in PF((devd[i] for i in range(len(paths))), depth=3))
    assert run_scenes(net, loader, out_b, pipelined=True) == len(paths)
    for p in paths:
        da, db = os.path.join(out_a, scene_key(p)), os.path.join(out_b, scene_key(p))
        for f in ("pred_class.npy", "pred_conf.npy", "pred_box.npy", "scene.npy"):
            assert np.array_equal(np.load(os.path.join(da, f)), np.load(os.path.join(db, f))), (p, f)
        ma, mb = pickle.load(open(os.path.join(da, "pred_mask"), "rb")), pickle.load(open(os.path.join(db, "pred_mask"), "rb"))
        assert len(ma) == len(mb) and all(np.array_equal(x, y) for x, y in zip(ma, mb))
        assert pickle.load(open(os.path.join(da, "pred_mask_index"), "rb")) == pickle.load(open(os.path.join(db, "pred_mask_index"), "rb"))
