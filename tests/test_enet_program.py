"""CPU: the BatchNorm-folded ENet layer program (lib/nets/enet_program.py, SURVEY row f2 host side) reproduces the
features of the unmodified reference ENet (tests/golden/enet_encoder.npz)."""
import numpy as np
import torch

from conftest import load_golden
from lib.nets import enet_program as E


def test_folded_program_matches_reference_features():
    g = load_golden("enet_encoder.npz")
    params = [torch.from_numpy(g[k]) for k in sorted(k for k in g if k.startswith("p"))]
    ops = E.compile_enet(params)
    kinds = [o[0] for o in ops]
    assert kinds.count("conv") == 1 + 2 * 3 + 4 * 3 + 12 * 3 + 4 * 4 and "affine_prelu" in kinds
    x = torch.from_numpy(np.random.default_rng(int(g["seed"])).standard_normal((1, 3, 256, 328)).astype(np.float32))
    with torch.no_grad():
        y = E.run_program(ops, x)
    ref = torch.from_numpy(g["features"])
    assert y.shape == ref.shape
    # folding re-associates the BatchNorm arithmetic: fp32 rounding only
    assert float((y - ref).abs().max()) < 2e-4 and float(((y - ref).norm() / ref.norm())) < 2e-6
