"""Drop-in test of the operator boundary: the reference's OWN Python -- network.py,
proposal_layer.py, pth_nms.py, roi_pool.py, unmodified, from git-ignored baseline/_ref/ -- runs on the GPU with its two cffi
extension modules rebound to libsis3d.so (`gpu_nms` -> sis3d_nms, `roi_pooling_forward_cuda` -> sis3d_roi_pool_fwd), in a
process of its own (oracle/run_reference_gpu.py).  Its predictions must equal (a) the CPU oracle, which is pinned to the
all-CPU run of the same files, and (b) this repo's Network.forward on the same inputs."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import sis3d_synth as synth
from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["odd_45x27x41", "cfg1_32"])
def test_reference_python_runs_over_libsis3d_extension_stubs(oracle, tag, tmp_path):
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "lib", "nets")):
        pytest.skip("baseline/_ref not staged (python __graft_entry__.py build, in the container that has /root/reference)")
    out = str(tmp_path / "ref_gpu.npz")
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_reference_gpu.py"), "--case", tag, "--out", out],
                       capture_output=True, text=True, timeout=900, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ref = dict(np.load(out))
    c = synth.CASES[tag]
    assert ref["convs_on_cuda"] == 1 and ref["gpu_nms_calls"] >= 1 and ref["roi_cuda_calls"] >= 1, "the rebound kernels must have run"
    ocfg, w, data, views = synth.build_case(oracle, c)
    want = oracle.forward(ocfg, w, data, views, fma_mode=1)
    # (a) reference-on-GPU over our kernels == CPU oracle (pinned to the all-CPU reference run)
    assert ref["rois"].shape == tuple(want["rois"].shape)
    np.testing.assert_allclose(ref["rois"], want["rois"].numpy(), atol=2e-3)
    assert np.array_equal(ref["level_inds"].reshape(-1), want["level_inds"].numpy().reshape(-1))
    assert np.array_equal(ref["cls_pred"], want["cls_pred"].numpy())
    np.testing.assert_allclose(ref["cls_prob"], want["cls_prob"].numpy(), atol=1e-4)
    # (b) == this repo's forward (default math mode)
    net, cfg = synth.make_net(c, keep_debug=False, math="exact")
    P = net.forward(synth.make_blobs(c, data, views), "TEST", None)
    torch.cuda.synchronize()
    np.testing.assert_allclose(P["rois"][0].cpu().numpy(), ref["rois"], atol=2e-3)
    assert np.array_equal(P["cls_pred"].cpu().numpy(), ref["cls_pred"])
    if c["use_mask"]:
        assert int(ref["n_masks"]) == len(P["mask_pred"][0]) == len(want["mask_pred"])
        for j, m in enumerate(P["mask_pred"][0]):
            assert float(np.abs(m.cpu().numpy() - ref[f"mask_{j}"]).max()) < 1e-3
