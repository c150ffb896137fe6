"""oracle/make_golden_vox2mesh.py -- TEST INFRASTRUCTURE ONLY; run in the build container: python oracle/make_golden_vox2mesh.py

Row f4 pin: executes the function definitions of the UNMODIFIED reference script tools/scannet_benchmark/vox2mesh.py (:23-121:
save_scannet_benchmark, load_pred, nn_search, export, load_matrix -- the text before its module-level argparse/main code, run
with exec in a scratch namespace; its `utils` import is only used by main) on seeded synthetic prediction folders and meshes,
and stores what they produce (the painted 400x200x400 scene as a sparse list, the benchmark text files) in
tests/golden/vox2mesh_reference.npz.  tests/test_vox2mesh_golden.py requires this repo's vectorised export to reproduce them."""
import os
import pickle
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SIS3D_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "vox2mesh_reference.npz")


def make_case(seed, dims=(60, 30, 50), n_box=9, n_vert=1500):
    """Seeded prediction folder contents + mesh vertices (shared with the test: same numpy generator calls)."""
    rng = np.random.default_rng(seed)
    box, masks = [], []
    for _ in range(n_box):
        lo = np.array([rng.integers(1, d - 10) for d in dims]) + rng.choice([0.0, 0.5, 0.49, -0.3], 3)
        size = rng.integers(2, 9, 3)
        hi = np.minimum(lo + size + rng.choice([0.0, 0.5, 0.51], 3), np.array(dims) - 1.6)
        lo = np.maximum(lo, 1)
        box.append(np.concatenate([lo, hi]))
        shape = (np.rint(hi).astype(int) - np.rint(lo).astype(int)).clip(0)
        masks.append((rng.random(tuple(shape)) < 0.6).astype(np.float32))
    box = np.array(box, dtype=np.float32)
    cls = rng.integers(1, 19, n_box)
    conf = rng.uniform(0.3, 0.99, n_box).astype(np.float32)
    keep = rng.random(n_box) < 0.75
    keep[0] = True
    w2g = np.eye(4)
    w2g[:3, :3] *= 1 / 0.046875
    w2g[:3, 3] = rng.uniform(1, 3, 3)
    verts = rng.uniform(0.02, 0.9, (n_vert, 3)) * (np.array(dims) * 0.046875)
    return dict(box=box, cls=cls, conf=conf, masks=masks, keep=keep, w2g=w2g, verts=verts, dims=dims)


def write_pred_folder(d, c):
    np.save(os.path.join(d, "pred_box.npy"), np.concatenate([c["box"], np.zeros((len(c["box"]), 1), np.float32)], 1))
    np.save(os.path.join(d, "pred_class.npy"), c["cls"])
    np.save(os.path.join(d, "pred_conf.npy"), c["conf"])
    pickle.dump([m for m, k in zip(c["masks"], c["keep"]) if k], open(os.path.join(d, "pred_mask"), "wb"))
    pickle.dump([bool(k) for k in c["keep"]], open(os.path.join(d, "pred_mask_index"), "wb"))


def reference_functions():
    src = open(os.path.join(REF, "tools", "scannet_benchmark", "vox2mesh.py")).read()
    head = src[:src.index("parser = argparse.ArgumentParser()")]
    sys.modules.setdefault("utils", types.ModuleType("utils"))
    ns = {}
    exec(compile(head, "reference:tools/scannet_benchmark/vox2mesh.py", "exec"), ns)
    return ns


def main():
    ns = reference_functions()
    g = {}
    for seed in (0, 1, 2):
        c = make_case(seed)
        with tempfile.TemporaryDirectory() as d:
            write_pred_folder(d, c)
            scene = ns["load_pred"](d)  # 400x200x400 float64, triple Python loop over box voxels
            out = os.path.join(d, "out")
            ns["export"](c["verts"], c["w2g"], scene, out, "scene0000_00")
            idx = np.argwhere(scene != 0)
            g[f"scene_idx_{seed}"] = idx.astype(np.int32)
            g[f"scene_val_{seed}"] = scene[scene != 0]
            lines = open(os.path.join(out, "scene0000_00.txt")).read().splitlines()
            g[f"lines_{seed}"] = np.array(lines)
            for ln in lines:
                f = ln.split()[0]
                g[f"mask_{seed}_{os.path.basename(f)}"] = np.loadtxt(os.path.join(out, f), dtype=np.uint8)
        p = os.path.join(tempfile.gettempdir(), "w2g.txt")
        open(p, "w").write("21.3 0 0 100.5\n0 21.3 0 50\n0 0 21.3 75.25\n0 0 0 1\n")
        g["load_matrix"] = ns["load_matrix"](p)
    np.savez_compressed(OUT, **g)
    print(OUT, os.path.getsize(OUT) // 1024, "kB;", {s: len(g[f"lines_{s}"]) for s in (0, 1, 2)}, "instances")


if __name__ == "__main__":
    main()
