import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("vapoursynth-mvtools_b200", "oracle", os.path.join("host", "minivs"), "tests", ""):
    q = os.path.join(ROOT, p)
    if q not in sys.path:
        sys.path.insert(0, q)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """The in-tree shared libraries; built once per session if missing (nvcc works without a GPU)."""
    import mvtools_b200 as mb
    need = [mb.LIB_PATH, mb.PLUGIN_PATH, os.path.join(ROOT, "oracle", "_ref", "liboracle.so"),
            os.path.join(ROOT, "host", "minivs", "libminivs.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    return True


@pytest.fixture(scope="session")
def have_ref(built):
    """True when the compiled reference (oracle/_ref/libmvtools_ref.so) is available."""
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libmvtools_ref.so"))


@pytest.fixture(scope="session")
def gpu_ctx(built):
    import mvtools_b200 as mb
    return mb.Context(0)
