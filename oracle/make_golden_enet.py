"""oracle/make_golden_enet.py -- TEST INFRASTRUCTURE ONLY; run in the build container:

    python oracle/make_golden_enet.py       # writes tests/golden/enet_encoder.npz

Row f2 groundwork: runs the UNMODIFIED reference ENet (lib/nets/enet.py: create_enet modules 0..25, the part
create_enet_for_3d keeps) through oracle/ref_harness.py with seeded random weights -- BatchNorm statistics, affine
terms and PReLU slopes randomised so that no layer is an identity -- on one seeded image, and stores the parameters (in
state_dict order), the input seed and the output features.  Pins oracle/port.py::enet_encoder.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "enet_encoder.npz")
SEED = 2024


def enet_image(seed, n=1):
    return np.random.default_rng(seed).standard_normal((n, 3, 256, 328)).astype(np.float32)


def main():
    rh.install()
    from lib.nets import enet  # the reference's module
    torch.manual_seed(SEED)
    model = enet.create_enet(21)
    enc = torch.nn.Sequential(*(model[i] for i in range(len(model) - 1))).eval()
    g = torch.Generator().manual_seed(SEED + 1)
    with torch.no_grad():
        for name, t in enc.state_dict().items():
            if name.endswith("running_mean"):
                t.copy_(torch.randn(t.shape, generator=g) * 0.2)
            elif name.endswith("running_var"):
                t.copy_(torch.rand(t.shape, generator=g) + 0.5)
        for m in enc.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            elif isinstance(m, torch.nn.PReLU):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5)
        x = torch.from_numpy(enet_image(SEED))
        y = enc(x)
    params = [v.detach().numpy().astype(np.float32) for k, v in enc.state_dict().items() if not k.endswith("num_batches_tracked")]
    out = {f"p{i:03d}": p for i, p in enumerate(params)}
    out.update(seed=np.array(SEED), features=y.numpy().astype(np.float32))
    np.savez_compressed(OUT, **out)
    print(f"{OUT}: {len(params)} parameter tensors, features {tuple(y.shape)}, {os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
