"""CPU: the native per-view constant builder (sis3d_view_params_host) is bit-identical to the reference's torch ops."""
import numpy as np
import torch

import sis3d_synth as synth
from lib.layer_utils import projection as P


def test_native_view_params_bit_identical_to_torch_ops():
    intr = synth.INTRINSIC_SCANNET.tolist()
    rng = np.random.default_rng(0)
    for trial in range(120):
        n = int(rng.integers(1, 9))
        dims = tuple(int(v) for v in rng.integers(16, 200, 3))
        v = synth.make_views(trial, (96, 48, 96), n, None)
        poses = v["poses"].copy()
        poses[:, :3, 3] += rng.normal(0, 1.0, (n, 3)).astype(np.float32)
        # random extra rotation so nothing is axis aligned
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        poses[:, :3, :3] = (poses[:, :3, :3] @ q.astype(np.float32))
        w2g = v["world2grid"].copy()
        w2g[:3, 3] = rng.normal(0, 20, 3).astype(np.float32)
        if trial % 3 == 0:
            w2g = np.stack([w2g] * n)  # one matrix per view
        a = P._view_params_impl(intr, (41, 32), 0.1, 4.0, dims, torch.from_numpy(poses), torch.from_numpy(w2g))
        b = P._view_params_torch(intr, (41, 32), 0.1, 4.0, dims, torch.from_numpy(poses), torch.from_numpy(w2g))
        assert torch.equal(a, b), f"trial {trial}: max diff {(a - b).abs().max()}"
