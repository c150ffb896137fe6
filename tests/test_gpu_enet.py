"""SURVEY row f2 on the GPU: the 2-D ENet encoder on the sis3d_enet_* kernels of libsis3d.so.
(a) features of the unmodified reference ENet (tests/golden/enet_encoder.npz, generator oracle/make_golden_enet.py);
(b) the whole TEST forward from RAW IMAGES (cfg.USE_IMAGES_GT=False, reference: lib/nets/network.py:204-205) against the CPU
    oracle (oracle.enet_encoder -> oracle.forward), integer outputs exact, masks within 1e-3, in the default math mode,
    through the synchronous forward (eager, then CUDA-graph replay) and the pipelined scene loop."""
import numpy as np
import pytest
import torch

import sis3d_synth as synth
from conftest import load_golden

pytestmark = pytest.mark.gpu


def _golden_params():
    g = load_golden("enet_encoder.npz")
    return g, [torch.from_numpy(g[k]) for k in sorted(k for k in g if k.startswith("p"))]


def test_enet_encoder_matches_reference_features():
    from lib.nets.enet import EnetEncoder
    g, params = _golden_params()
    x = torch.from_numpy(np.random.default_rng(int(g["seed"])).standard_normal((1, 3, 256, 328)).astype(np.float32)).cuda()
    enc = EnetEncoder(params, "cuda:0")
    y = enc(x)
    torch.cuda.synchronize()
    ref = torch.from_numpy(g["features"]).cuda()
    assert y.shape == ref.shape == (1, 128, 32, 41)
    torch.testing.assert_close(y, ref, atol=2e-5, rtol=1e-5)
    y5 = enc(x.repeat(5, 1, 1, 1))  # batch of views: every image independent
    torch.cuda.synchronize()
    assert all(torch.equal(y5[i], y[0]) for i in range(5))


def _net_from_images(c, params, math="exact"):
    from lib.utils.config import cfg
    net, _ = synth.make_net(c, keep_debug=False, math=math)  # builds the 3-D weights; rebuild with the encoder declared
    sd3d = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    del net
    cfg.USE_IMAGES_GT = False
    from lib.nets import backbones
    net = getattr(backbones, cfg.NET)()
    net.init_modules()
    sd = net.state_dict()
    names = [n for n in net._enet_names]
    assert len(names) == len(params)
    for n, p in zip(names, params):
        assert tuple(sd[n].shape) == tuple(p.shape), n
        sd[n] = p
    sd.update(sd3d)
    net.load_state_dict(sd, strict=True)
    net._keep_debug = False
    net.set_conv_math(math)
    return net, cfg


def test_forward_from_raw_images_matches_oracle(oracle):
    c = synth.CASES["odd_45x27x41"]
    g, params = _golden_params()
    ocfg, w, data, views = synth.build_case(oracle, c)
    images = np.random.default_rng(77).standard_normal((c["n_img"], 3, 256, 328)).astype(np.float32)
    with torch.no_grad():
        feats = oracle.enet_encoder(params, torch.from_numpy(images)).numpy()
    v2 = dict(views, feats=feats)
    want = oracle.forward(ocfg, w, data, v2, fma_mode=1)
    net, cfg = _net_from_images(c, params)
    blobs = synth.make_blobs(c, data, dict(views, feats=images))

    def check(P):
        got, ref = P["rois"][0].cpu().numpy(), want["rois"].numpy()
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, atol=2e-3)
        assert np.array_equal(P["level_inds"][0].cpu().numpy(), want["level_inds"].numpy())
        assert np.array_equal(P["cls_pred"].cpu().numpy(), want["cls_pred"].numpy())
        det = P["detections_host"]
        assert np.array_equal(det[:, 8] > 0.5, want["mask_keep"])
        assert np.array_equal(det[:, 9:15].astype(np.int64), want["mask_crops"])
        assert len(P["mask_pred"][0]) == len(want["mask_pred"])
        for m, r in zip(P["mask_pred"][0], want["mask_pred"]):
            assert float((m.cpu() - r).abs().max()) < 1e-3
    for _ in range(3):  # eager first sight, then capture + replay with the encoder inside the graph
        check(net.forward(blobs, "TEST", None))
    n = 0
    for _, P in net.forward_pipelined(iter([blobs] * 8)):
        check(P)
        n += 1
    assert n == 8
