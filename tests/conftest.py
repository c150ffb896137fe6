import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "3d-sis_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (test infrastructure): oracle/port.py with liboracle.so built on demand."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    from oracle import port
    return port


def load_golden(name):
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def emu_lib(tmp_path_factory):
    """TEST INFRASTRUCTURE: the shipped CUDA-core / proposal / RoI / projection / sparse sources compiled for the HOST
    (tools/cuda_host_emu.py; everything except the TMA / tcgen05 kernels of conv_tc.cu) as one library with libsis3d's
    C entry points operating on host memory."""
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import cuda_host_emu
    out = str(tmp_path_factory.mktemp("emu") / "libsis3d_emu.so")
    # SIS3D_EMU_TSAN=1 + LD_PRELOAD=libtsan.so turns the emulated tests into a shared-memory racecheck (profiles/)
    cuda_host_emu.build(out, [os.path.join(ROOT, "3d-sis_b200", "csrc", f)
                              for f in ("conv_simt.cu", "rpn.cu", "roi.cu", "sparse.cu", "project.cu", "io.cu")],
                        tsan=os.environ.get("SIS3D_EMU_TSAN") == "1")
    lib = C.CDLL(out)
    for f in ("sis3d_nms_workspace_bytes", "sis3d_rpn_workspace_bytes", "sis3d_project_compact_workspace_bytes",
              "sis3d_backproject_conv_k2s2_workspace_bytes", "sis3d_linear_workspace_bytes"):
        getattr(lib, f).restype = C.c_size_t
    lib.sis3d_strerror.restype = C.c_char_p
    return lib


@pytest.fixture
def host_S(monkeypatch, emu_lib):
    """Point the ctypes layer (lib._sis3d) at the emulated library and let it accept host tensors, so that the Python
    wrappers above the C ABI run unchanged on the CPU."""
    import ctypes as C
    from lib import _sis3d as S
    monkeypatch.setattr(S, "lib", emu_lib)
    monkeypatch.setattr(S, "ptr", lambda t: None if t is None else C.c_void_p(t.data_ptr()))
    monkeypatch.setattr(S, "stream", lambda: None)
    return S
