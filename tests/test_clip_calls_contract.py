"""CPU test (no GPU) of the CONTRACT of the clip-level entry points (mvb200_*_clip_host: which source frames a window of
output frames needs, how frames at scene changes / clip ends are made, what the usable flags say): the bodies of the GPU
tests in test_gpu_fused.py / test_gpu_flow2.py run unchanged against the stand-in engine of tests/mock (which forwards
to the oracle).  It pins the expectations the GPU tests hold the CUDA engine to -- windows, halos, golden indexing --
so that a failure of those tests on a GPU points at the engine and not at the test.  Test infrastructure only."""
import ctypes as C

import pytest

import mvtools_b200 as mb
from test_plugin_glue import mock_plugin  # noqa: F401  (fixture: builds tests/mock/libmvtools_mockengine.so)

CLIP_CALLS = ("mvb200_flow_clip_host", "mvb200_flowinter_clip_host", "mvb200_flowblur_clip_host", "mvb200_blockfps_clip_host", "mvb200_flowfps_clip_host", "mvb200_compensate_clip_host")


class _Ctx:
    handle = None


@pytest.fixture()
def stand_in(mock_plugin, monkeypatch):  # noqa: F811
    real = mb.lib()
    mock = C.CDLL(mock_plugin)
    for name in CLIP_CALLS:
        getattr(mock, name).argtypes = getattr(real, name).argtypes

    class Lib:
        def __getattr__(self, k):
            return getattr(mock, k) if k in CLIP_CALLS else getattr(real, k)

    monkeypatch.setattr(mb, "lib", lambda: Lib())
    return _Ctx()


@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "noise_badblocks_blk8"])
def test_clip_call_expectations_hold_on_the_stand_in_engine(stand_in, case):
    import test_gpu_flow2
    import test_gpu_fused
    test_gpu_fused.test_flow_clip_host_matches_goldens(stand_in, case)
    test_gpu_fused.test_flowinter_clip_host_matches_goldens(stand_in, case)
    test_gpu_fused.test_flowblur_clip_host_matches_goldens(stand_in, case)
    test_gpu_fused.test_blockfps_clip_host_matches_goldens(stand_in, case)
    test_gpu_flow2.test_flowfps_clip_host_matches_goldens(stand_in, case)
    import pipeline
    if case in pipeline.EXTRA:
        test_gpu_flow2.test_compensate_clip_host_matches_goldens(stand_in, case)
