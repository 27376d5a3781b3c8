"""CPU tests of the host side of the C ABI: the library loads, exports every symbol the
header declares, derives parameters like the reference's *Create functions and reports the
reference's error messages."""
import ctypes as C
import os
import re

import pytest

import mvtools_b200 as mb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "mvtools_b200.h")).read()
    names = re.findall(r"MVB200_API\s+[\w\s\*]+?\b(mvb200_\w+)\s*\(", hdr)
    assert len(names) >= 20
    L = C.CDLL(mb.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libmvtools_b200.so does not export %s" % n
    assert b"mvtools_b200" in mb.lib().mvb200_version()


def _exported(path):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", path], stdout=subprocess.PIPE, text=True, check=True).stdout
    return sorted(line.split()[-1] for line in out.splitlines() if line.strip())


def test_plugin_exports_entry_point(built):
    """exactly one export, like the reference module (EntryPoint.c:28-29; -fvisibility=hidden, meson.build:15): no libstdc++
    or template symbols leak into a host that loads other C++ plugins"""
    L = C.CDLL(mb.PLUGIN_PATH)
    assert hasattr(L, "VapourSynthPluginInit2")
    assert _exported(mb.PLUGIN_PATH) == ["VapourSynthPluginInit2"]


def test_engine_exports_only_the_c_abi(built):
    names = _exported(mb.LIB_PATH)
    assert names and all(n.startswith("mvb200_") for n in names), [n for n in names if not n.startswith("mvb200_")][:5]


def test_super_geometry_matches_survey_table(built):
    # super frame sizes for the BASELINE configs
    for (w, h, pel, sw, sh, lv) in ((640, 360, 1, 672, 978, 8), (1920, 1080, 2, 1952, 5822, 10), (3840, 2160, 2, 3872, 11254, 11)):
        sp = mb.super_params(w, h, 8, 1, 1, pel=pel)
        assert (sp.super_width, sp.super_height, sp.levels) == (sw, sh, lv)


def test_analyse_geometry_matches_survey_table(built):
    sp = mb.super_params(1920, 1080, 8, 1, 1, pel=2)
    ap = mb.analyse_params(sp, 10, blksize=8, overlap=4)
    assert (ap.ad.nBlkX, ap.ad.nBlkY, ap.ad.nLvCount) == (479, 269, 8)
    assert [ap.nBlkX[i] * ap.nBlkY[i] for i in range(8)] == [128851, 32026, 7854, 1888, 435, 98, 18, 2]
    ap = mb.analyse_params(sp, 10, blksize=16, overlap=8)
    assert (ap.ad.nBlkX, ap.ad.nBlkY, ap.ad.nLvCount) == (239, 134, 7)
    # truemotion defaults, MVAnalyse.c:303-330,477-483
    assert (ap.nLambda, ap.lsad, ap.pnew, ap.plevel, ap.global_, ap.badSAD) == (4000, 4800, 50, 1, 1, 40000)
    sp16 = mb.super_params(1920, 1080, 16, 1, 1, pel=2)
    ap = mb.analyse_params(sp16, 10, blksize=16, overlap=8)
    assert (ap.nLambda, ap.lsad, ap.badSAD) == (1028000, 1233600, 10280000)


@pytest.mark.parametrize("kw,msg", [
    (dict(search=9), "Analyse: search must be between 0 and 7 (inclusive)."),
    (dict(blksize=12), "Analyse: the block size must be 4x4, 8x4, 8x8, 16x2, 16x8, 16x16, 32x16, 32x32, 64x32, 64x64, 128x64, or 128x128."),
    (dict(blksize=8, overlap=6), "Analyse: overlap must be at most half of blksize, overlapv must be at most half of blksizev, and they both need to be at least 0."),
    (dict(blksize=8, overlap=3), "Analyse: The requested overlap is incompatible with the super clip's subsampling."),
    (dict(dct=11), "Analyse: dct must be between 0 and 10 (inclusive)."),
    (dict(pnew=300), "Analyse: pnew must be between 0 and 256 (inclusive)."),
    (dict(levels=40), "Analyse: invalid number of levels."),
    (dict(blksize=4, divide=1), "Analyse: blksize and blksizev must be at least 8 when divide=True."),
])
def test_analyse_errors_are_the_references(built, kw, msg):
    sp = mb.super_params(640, 360, 8, 1, 1)
    with pytest.raises(mb.Error) as e:
        mb.analyse_params(sp, 10, **kw)
    assert str(e.value) == msg


def test_super_errors(built):
    with pytest.raises(mb.Error) as e:
        mb.super_params(640, 360, 8, 1, 1, pel=3)
    assert str(e.value) == "Super: pel must be 1, 2, or 4."


def test_super_pelclip_check(built):
    """mvsuperCreate's pelclip rules (MVSuper.c:229-253)"""
    sp = mb.super_params(160, 96, 8, 1, 1, pel=2, hpad=8, vpad=8)
    assert mb.super_pelclip_check(sp, 320, 192) == (True, False)
    assert mb.super_pelclip_check(sp, (160 + 16) * 2, (96 + 16) * 2) == (True, True)
    with pytest.raises(mb.Error) as e:
        mb.super_pelclip_check(sp, 321, 192)
    assert str(e.value) == "Super: pelclip's dimensions must be multiples of the input clip's dimensions."
    with pytest.raises(mb.Error) as e:
        mb.super_pelclip_check(sp, 320, 192, same_format=False)
    assert str(e.value) == "Super: pelclip must have the same format as the input clip, and it must have constant dimensions."
    sp1 = mb.super_params(160, 96, 8, 1, 1, pel=1)
    assert mb.super_pelclip_check(sp1, 123, 45) == (False, False)  # ignored for pel 1


def test_degrain_errors_and_thresholds(built):
    sp = mb.super_params(640, 360, 8, 1, 1)
    bw = mb.analyse_params(sp, 10, isb=1, blksize=16, overlap=8)
    fw = mb.analyse_params(sp, 10, isb=0, blksize=16, overlap=8)
    dp = mb.degrain_params(sp, [bw.ad, fw.ad], 640, 360, 1)
    # scaleThSCD (MVAnalysisData.c:7-31): 400*256/64 = 1600, +1600/4*2 = 2400; thsad 400 -> 2400
    assert (dp.nSCD1, dp.nSCD2, dp.thSAD[0], dp.thSAD[1]) == (2400, 130 * bw.ad.nBlkX * bw.ad.nBlkY // 256, 2400, 2400)
    with pytest.raises(mb.Error) as e:
        mb.degrain_params(sp, [fw.ad, bw.ad], 640, 360, 1)
    assert str(e.value) == "Degrain1: mvfw must be generated with isb=False."
    with pytest.raises(mb.Error) as e:
        mb.degrain_params(sp, [bw.ad, fw.ad], 640, 360, 1, thscd1=20000)
    assert str(e.value) == "Degrain1: thscd1 can be at most 16320."
    with pytest.raises(mb.Error) as e:
        mb.degrain_params(sp, [bw.ad, fw.ad], 320, 360, 1)
    assert str(e.value) == "Degrain1: wrong source or super clip frame size."


def test_compute_fails_loudly_without_gpu(built):
    """No CPU fallback: on a box without a CUDA device the context cannot be created."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mb.Error) as e:
        mb.Context(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_window_plan_query(built):
    """The window sizing the plugin's fused paths and bench.py share (host arithmetic only, no device needed)."""
    import ctypes as C
    L = mb.lib()

    def plan(w, h, bits, blk, ov, nvec, lookahead=-1, contexts=2, budget=0, devmem=180 << 30, ndev=1):
        sp = mb.super_params(w, h, bits, 1, 1, pel=2)
        ap = mb.analyse_params(sp, 1000, blksize=blk, overlap=ov)
        cl = mb.clip_layout(sp)
        q, wp = mb.SearchPlan(), mb.WindowPlan()
        assert L.mvb200_search_plan_query(C.byref(ap), C.byref(sp), 0, C.byref(q)) == 0
        assert L.mvb200_window_plan_query(C.byref(ap), C.byref(sp), nvec, lookahead, cl.frame_bytes, cl.frame_bytes, ndev, contexts, budget, devmem, C.byref(wp)) == 0
        return q, wp, cl

    q, wp, cl = plan(1920, 1080, 8, 16, 8, 6, budget=32 << 30)
    assert wp.frames == -(-q.chains_per_wave // 6) == 296 and wp.contexts_per_device == 2  # the headline: one wave of chain slots
    q, wp, cl = plan(1920, 1080, 8, 16, 8, 1, budget=32 << 30)
    assert wp.frames == 768  # a single vector clip: bounded latency
    q, wp, cl = plan(1920, 1080, 8, 16, 8, 6, lookahead=40)
    assert wp.frames == 40  # an explicit look-ahead is taken as given
    q, wp, cl = plan(1920, 1080, 8, 16, 8, 6, budget=2 << 30)
    assert wp.frames == (2 << 30) // (4 * 2 * cl.frame_bytes) and wp.pinned_budget == 2 << 30  # the pinned windows fit the budget
    q, wp8, cl = plan(1920, 1080, 8, 16, 8, 6, budget=2 << 30, ndev=8)
    assert wp8.frames < wp.frames  # more contexts, more windows in flight
    q, wp, cl = plan(3840, 2160, 16, 16, 8, 6, budget=32 << 30, devmem=40 << 30)
    assert wp.contexts_per_device == 2 and 16 <= wp.frames < 100  # a small device: the contexts that were asked for, shorter windows
    q, wp, cl = plan(3840, 2160, 16, 16, 8, 6, budget=32 << 30, devmem=4 << 30)
    assert wp.contexts_per_device == 1 and wp.frames == 16  # ... that cannot even hold two minimal windows: one context
    # the number of contexts left to the plan (0): two waves of chains in flight per device are the target
    q, wp, cl = plan(1920, 1080, 8, 16, 8, 6, contexts=0, budget=32 << 30)
    assert (wp.frames, wp.contexts_per_device) == (296, 2)  # whole waves: two contexts
    q, wp4, cl = plan(3840, 2160, 8, 8, 0, 1, contexts=0, budget=24 << 30)
    q, wp1, cl = plan(3840, 2160, 8, 8, 0, 1, contexts=1, budget=24 << 30)
    # 4K, one vector clip per frame: a window is what the pinned budget allows, far from a wave -- more, smaller windows
    assert wp4.contexts_per_device == 4 and wp4.frames < wp1.frames and 4 * wp4.frames > 1.8 * wp1.frames
    assert (4 + 2) * wp4.frames * 2 * cl.frame_bytes <= 24 << 30
    q, wp, cl = plan(1920, 1080, 8, 16, 8, 6)
    assert wp.pinned_budget >= 2 << 30 and wp.pinned_budget <= 32 << 30
