"""CPU: the ENet encoder restatement (oracle/port.py::enet_encoder) against features produced by the
unmodified reference ENet (tests/golden/enet_encoder.npz, generator: oracle/make_golden_enet.py)."""
import numpy as np
import torch

from conftest import load_golden


def test_enet_encoder_port_matches_reference_features(oracle):
    g = load_golden("enet_encoder.npz")
    params = [torch.from_numpy(g[k]) for k in sorted(k for k in g if k.startswith("p"))]
    x = torch.from_numpy(np.random.default_rng(int(g["seed"])).standard_normal((1, 3, 256, 328)).astype(np.float32))
    with torch.no_grad():
        y = oracle.enet_encoder(params, x)
    ref = torch.from_numpy(g["features"])
    assert y.shape == ref.shape == (1, 128, 32, 41)
    assert float(ref.abs().mean()) > 1e-2  # a live signal, not a collapsed network
    torch.testing.assert_close(y, ref, atol=1e-5, rtol=1e-5)


def test_enet_encoder_rejects_wrong_parameter_count(oracle):
    g = load_golden("enet_encoder.npz")
    params = [torch.from_numpy(g[k]) for k in sorted(k for k in g if k.startswith("p"))]
    x = torch.zeros(1, 3, 64, 64)
    try:
        oracle.enet_encoder(params + [torch.zeros(1)], x)
    except ValueError:
        return
    raise AssertionError("extra parameters must be rejected")
