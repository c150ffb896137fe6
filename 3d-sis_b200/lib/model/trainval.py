"""Inference drivers with the reference's entry points `test(args)` / `benchmark(args)`
(lib/model/trainval.py:75-93, SolverWrapper.test 769-941, .benchmark 634-767).

Per scene: forward (the mask head runs ONCE -- the reference runs it in forward and again in the
driver, trainval.py:882-897), then the on-disk results the ScanNet tooling consumes
(trainval.py:839-845, 909-911): pred_class.npy, pred_conf.npy, pred_box.npy, scene.npy, pickled pred_mask /
pred_mask_index.  Class-specific decode, clipping, confidence and degenerate-box filtering were already
done on the device (csrc/roi.cu detect_decode_kernel) and arrive in one [n,16] table.
Training (`train`) is out of scope of the B200 inference path.
"""
from __future__ import annotations

import os
import pickle
import time

import numpy as np
import torch

from lib.datasets.dataset import Dataset, collate_fn
from lib.utils.config import cfg


def scene_key(scene_id):
    return os.path.basename(str(scene_id))[:12]


def detections_from_predictions(net):
    """-> pred_class int64[n], pred_conf float64[n], pred_box float32[n,6], keep bool[n] (trainval.py:825-858)."""
    det = net._predictions["detections_host"]
    return det[:, 7].astype(np.int64), det[:, 6].astype(np.float64), det[:, 0:6].astype(np.float32), det[:, 8] > 0.5


def save_scene_results(out_dir, scene_id, blobs, net):
    d = os.path.join(out_dir, scene_key(scene_id))
    os.makedirs(d, exist_ok=True)
    pred_class, pred_conf, pred_box, keep = detections_from_predictions(net)
    np.save(os.path.join(d, "pred_class"), pred_class)
    np.save(os.path.join(d, "pred_conf"), pred_conf)
    np.save(os.path.join(d, "pred_box"), pred_box)
    np.save(os.path.join(d, "scene"), np.where(blobs["data"][0, 0].numpy() <= 1, 1, 0))
    if cfg.USE_MASK:
        masks = []
        if keep.any():  # one D2H of the thresholded predicted-class channels (csrc/roi.cu mask_select_kernel)
            bits = net._predictions["mask_bits"].cpu().numpy()
            offs, sizes = net._predictions["mask_offsets"], net._predictions["mask_sizes"]
            for j in range(len(sizes)):
                masks.append(bits[int(offs[j]):int(offs[j + 1])].reshape(tuple(int(v) for v in sizes[j])).astype(np.float32))
        with open(os.path.join(d, "pred_mask"), "wb") as f:
            pickle.dump(masks, f)
        with open(os.path.join(d, "pred_mask_index"), "wb") as f:
            pickle.dump([bool(k) for k in keep], f)
    return d


def run_scenes(net, data_loader, out_dir, skip_existing=False):
    os.makedirs(out_dir, exist_ok=True)
    t0 = time.time()
    done = 0
    for blobs in data_loader:
        if skip_existing and os.path.isdir(os.path.join(out_dir, scene_key(blobs["id"][0]))):  # trainval.py:647-653
            continue
        net.forward(blobs, "TEST", None)
        save_scene_results(out_dir, blobs["id"][0], blobs, net)
        done += 1
    torch.cuda.synchronize()
    print("It took {:.3f}s for test on whole scenes".format(time.time() - t0))
    return done


def _build(args, mode, view_provider=None):
    from lib.nets import backbones
    filelist = cfg.TEST_FILELIST
    dataset = Dataset(filelist, mode, view_provider=view_provider)
    loader = torch.utils.data.DataLoader(dataset, batch_size=1, shuffle=False, num_workers=getattr(args, "num_workers", 0),
                                         collate_fn=collate_fn)
    net = getattr(backbones, cfg.NET)()
    net.init_modules()
    ckpt = os.path.join(args.output_dir, "step_{}.pth".format(args.step))
    net.load_state_dict(torch.load(ckpt, map_location="cpu"), strict=False)
    return net, loader


def test(args, view_provider=None):
    net, loader = _build(args, "test", view_provider)
    return run_scenes(net, loader, cfg.TEST_SAVE_DIR)


def benchmark(args, view_provider=None):
    net, loader = _build(args, "benchmark", view_provider)
    return run_scenes(net, loader, cfg.TEST_SAVE_DIR, skip_existing=True)


def train(args):
    raise NotImplementedError("training is out of scope of the B200 inference hot path (SURVEY section 8)")
