"""oracle/make_golden_dataset.py -- TEST INFRASTRUCTURE ONLY; run in the build container:  python oracle/make_golden_dataset.py

Row f3 pin: a small synthetic `.chunk` file + frame folder (depth PNGs, colour JPEGs, pose files, label csv) is committed
under tests/golden/dataset/, and the UNMODIFIED reference reader (lib/datasets/dataset.py:45-218 `Dataset.__getitem__` with
BinaryReader.py:10-36, through oracle/ref_harness.py) parses those exact bytes; its item dict is stored as
tests/golden/dataset/reference_item.npz.  tests/test_dataset_golden.py then requires this repo's reader
(lib/datasets/{scene_io,dataset,frames}.py) to return the same arrays from the same files.  The container bytes come from this
repo's writer (the original writer is a Windows C++ tool, datagen/SceneSampler/main.cpp:348-395): if they were not in the
reference's format the reference reader would fail or disagree here."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
OUT = os.path.join(ROOT, "tests", "golden", "dataset")


def build_files():
    """Everything the reader consumes, generated from a fixed seed with numpy + PIL only (no repo code except the writer)."""
    import importlib.util
    from PIL import Image
    spec = importlib.util.spec_from_file_location("scene_io_w", os.path.join(ROOT, "3d-sis_b200", "lib", "datasets", "scene_io.py"))
    sio = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sio)
    rng = np.random.default_rng(31)
    os.makedirs(OUT, exist_ok=True)
    X, Y, Z = 24, 52, 20  # Y > 48: exercises the max-height crop
    sdf = rng.normal(0, 2.5, (X, Y, Z)).astype(np.float32)
    # (min xyz, max xyz, raw nyu40 label): fractional corners (floor/ceil), one box sticking out of the 96x48x96 volume
    # (part-in-volume < 1 -> dropped by KEEP_THRESH = 1), one of a zero-weight class, one taller than the height crop
    boxes = np.array([[2.3, 1.2, 3.7, 10.6, 9.1, 12.2, 3], [5.5, 0.0, 1.5, 20.2, 30.9, 18.4, 4], [-3.2, 2.0, 2.0, 6.0, 8.0, 9.0, 5],
                      [1.0, 1.0, 1.0, 7.5, 6.5, 5.5, 1], [3.0, 2.0, 4.0, 9.0, 50.5, 11.0, 3]], dtype=np.float32)
    masks = []
    for b in boxes:
        d = np.maximum(np.ceil(b[3:6]) - np.floor(b[0:3]), 1).astype(int)
        m = rng.integers(0, 4, tuple(d)).astype(np.uint16)  # values > 1 are cleared by the reader
        masks.append((int(b[6]), m))
    w2g = np.array([[21.3, 0, 0, -10.5], [0, 21.3, 0, -3.25], [0, 0, 21.3, 7.75], [0, 0, 0, 1]], dtype=np.float64)
    chunk = os.path.join(OUT, "sample__0.chunk")
    sio.write_scene(chunk, sdf, boxes, masks, part_in_volume=np.ones(len(boxes), np.float32), world2grid=w2g, frame_ids=[0, 20])
    base = os.path.join(OUT, "frames_square")
    for sub in ("depth", "color", "pose"):
        os.makedirs(os.path.join(base, "sample", sub), exist_ok=True)
    for fid in (0, 20):
        depth = rng.integers(300, 4200, (60, 80)).astype(np.uint16)      # millimetres, 80x60 -> nearest resize + centre crop
        Image.fromarray(depth).save(os.path.join(base, "sample", "depth", f"{fid}.png"))
        col = rng.integers(0, 256, (128, 170, 3)).astype(np.uint8)       # 170x128 -> 328x256
        Image.fromarray(col).save(os.path.join(base, "sample", "color", f"{fid}.jpg"), quality=92)
        pose = np.eye(4) + rng.normal(0, 0.3, (4, 4))
        with open(os.path.join(base, "sample", "pose", f"{fid}.txt"), "w") as f:
            for r in pose:
                f.write(" ".join(f"{v:.6f}" for v in r) + "\n")
    with open(os.path.join(OUT, "labels.csv"), "w") as f:
        f.write("nyu40id,nyu40class,mappedId,mappedIdConsecutive,weight\n1,wall,(ignore),4,0.0\n3,cabinet,3,1,3.96\n"
                "4,bed,4,2,5.45\n5,chair,5,3,1.25\n")
    return chunk, base


def main():
    chunk, base = build_files()
    import torchvision.transforms  # noqa: F401  (before the harness stubs absent deps: torch._dynamo probes module specs)
    import ref_harness as rh
    import types
    from PIL import Image
    misc = types.ModuleType("scipy.misc")
    misc.imread = lambda f: np.array(Image.open(f))  # scipy.misc.imread (removed from scipy) was a thin PIL wrapper
    import scipy
    sys.modules["scipy.misc"] = misc
    scipy.misc = misc
    cfg = rh.load_cfg("ScanNet/rpn_class_mask_5.yml", USE_IMAGES=True, USE_IMAGES_GT=False, USE_MASK=True,
                      BASE_IMAGE_PATH=base, LABEL_MAP=os.path.join(OUT, "labels.csv"))
    from lib.datasets.dataset import Dataset
    lst = os.path.join(OUT, "filelist.txt")
    with open(lst, "w") as f:
        f.write(chunk + "\n")
    item = Dataset(lst, "chunk")[0]
    os.remove(lst)
    v = item["nearest_images"]
    g = dict(data=item["data"], gt_box=np.asarray(item["gt_box"], dtype=np.float32), n_mask=np.array(len(item["gt_mask"])),
             depths=np.stack(v["depths"]), images=np.stack([np.asarray(i) for i in v["images"]]), poses=np.stack(v["poses"]),
             world2grid=np.asarray(v["world2grid"]), frameids=np.asarray([int(i) for i in v["frameids"]]),
             keep_thresh=np.array(float(cfg.KEEP_THRESH)))
    for j, m in enumerate(item["gt_mask"]):
        g[f"mask_{j}"] = m
    np.savez_compressed(os.path.join(OUT, "reference_item.npz"), **g)
    print("reference reader:", {k: getattr(x, "shape", x) for k, x in g.items()})


if __name__ == "__main__":
    main()
