"""CPU: SURVEY row f4 pinned -- the vectorised voxel -> mesh export against what the UNMODIFIED reference script
(tools/scannet_benchmark/vox2mesh.py:23-121, executed by oracle/make_golden_vox2mesh.py) produced on the same seeded
prediction folders and meshes: the painted 400x200x400 scene (first-come instance painting, banker's rounding of box
corners), the nearest-neighbour vertex lookup, instance order, and the benchmark text files byte for byte."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden
from tools.scannet_benchmark import vox2mesh as V

sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden_vox2mesh import make_case, write_pred_folder  # noqa: E402  (the seeded inputs; no reference code involved)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_export_equals_reference_script(seed, tmp_path):
    g = load_golden("vox2mesh_reference.npz")
    c = make_case(seed)
    pred = tmp_path / "pred"
    pred.mkdir()
    write_pred_folder(str(pred), c)
    scene = V.load_pred(str(pred))  # default dims: the reference's fixed 400x200x400 canvas
    assert scene.shape == (400, 200, 400)
    idx = np.argwhere(scene != 0)
    assert np.array_equal(idx, g[f"scene_idx_{seed}"])
    assert np.array_equal(scene[scene != 0], g[f"scene_val_{seed}"])  # float64 values: box_ind*100 + class + conf - 0.01
    out = tmp_path / "out"
    V.export(c["verts"], c["w2g"], scene, str(out), "scene0000_00")
    lines = open(out / "scene0000_00.txt").read().splitlines()
    assert lines == [str(s) for s in g[f"lines_{seed}"]]  # same instances, same order, same class / score text
    for ln in lines:
        f = ln.split()[0]
        assert np.array_equal(np.loadtxt(out / f, dtype=np.uint8), g[f"mask_{seed}_{os.path.basename(f)}"])


def test_load_matrix_equals_reference(tmp_path):
    p = tmp_path / "w2g.txt"
    p.write_text("21.3 0 0 100.5\n0 21.3 0 50\n0 0 21.3 75.25\n0 0 0 1\n")
    assert np.array_equal(V.load_matrix(str(p)), load_golden("vox2mesh_reference.npz")["load_matrix"])
