"""CPU test (no GPU) of the VapourSynth plugin's host-side glue: tests/mock builds the product's plugin.cpp, unchanged,
against a stand-in engine whose per-frame entry points forward to the CPU oracle (tests/mock/mock_engine.cpp).  A script
driven through that plugin by the mini host reproduces the reference's goldens only if the glue requests the right
frames, picks the right reference frames / blobs / field shifts, attaches the right properties and takes the
reference's fall-back paths at scene changes.  Test infrastructure only: the product plugin links libmvtools_b200.so
(CUDA) and has no CPU path; the GPU tests in test_plugin.py drive the real one."""
import json
import os
import subprocess
import sys

import pytest

import pipeline
import pipeline_flow

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))
GOLD_FLOW = json.load(open(os.path.join(HERE, "golden", "golden_flow.json")))


@pytest.fixture(scope="module")
def mock_plugin(built):
    d = os.path.join(HERE, "mock")
    so = os.path.join(d, "libmvtools_mockengine.so")
    pkg = os.path.join(ROOT, "vapoursynth-mvtools_b200")
    srcs = [os.path.join(pkg, "plugin", "plugin.cpp"), os.path.join(d, "mock_engine.cpp"), os.path.join(pkg, "csrc", "host_logic.cpp")]
    deps = srcs + [os.path.join(ROOT, "include", "mvtools_b200.h"), os.path.join(ROOT, "oracle", "_ref", "liboracle.so")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-fvisibility=hidden", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "include", "vs"),
                               "-shared", "-Wl,-Bsymbolic", "-o", so] + srcs + ["-L" + os.path.join(ROOT, "oracle", "_ref"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_ref")])
    return so


@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3", "divide2_blk16_esa_neg_badrange", "pelclip2_p8", "noise_badblocks_blk8", "blk64_420_ov32_r2", "blk32x16_penalties_levels2", "p8_440_blk8_ov4_r2"])
def test_hot_path_glue(mock_plugin, case):
    got = pipeline.run_reference(case, plugin_path=mock_plugin)
    for stage in ("super", "vectors", "degrain", "adata", "recalc", "compensate", "recalc_adata"):
        if stage in got:
            assert got[stage] == GOLD[case][stage], "%s: %s" % (case, stage)


@pytest.mark.parametrize("case", sorted(pipeline_flow.FLOW))
def test_flow_group_glue(mock_plugin, case):
    got = pipeline_flow.run_reference_flow(case, plugin_path=mock_plugin)
    for stage in ("scd", "mask", "flow", "finest", "inter", "fps", "blur", "bfps"):
        assert got[stage] == GOLD_FLOW[case][stage], "%s: %s" % (case, stage)


@pytest.mark.parametrize("window", [2, 3, 64])
@pytest.mark.parametrize("case", ["blk16_ov8_pel2_p8_r3", "divide2_blk16_esa_neg_badrange", "noise_badblocks_blk8"])
def test_analyse_lookahead_glue(mock_plugin, monkeypatch, case, window):
    """MVB200_LOOKAHEAD: mv.Analyse searches a window of frames per engine call and serves the rest from its cache;
    the vectors (and everything computed from them) must not change, whatever the window and the request order"""
    monkeypatch.setenv("MVB200_LOOKAHEAD", str(window))
    got = pipeline.run_reference(case, plugin_path=mock_plugin)
    for stage in ("super", "vectors", "degrain", "adata", "recalc", "compensate"):
        if stage in got:
            assert got[stage] == GOLD[case][stage], "%s (window %d): %s" % (case, window, stage)


@pytest.mark.parametrize("window", [2, 5])
@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3", "gray_blk4_pel1_rf0", "p8_422_blk8x4_sharp0_rf1", "noise_badblocks_blk8", "c1_640x360_p8_blk8_pel1", "r6_blk8_ov4_p8", "r5_blk16_p16_422",
                                  "r4_gray_blk8x4", "p8_440_blk8_ov4_r2", "super_nochroma_blk8_r2"])
def test_degrain_fused_path_glue(mock_plugin, monkeypatch, case, window):
    """MVB200_LOOKAHEAD: a DegrainN whose inputs are this plugin's own Super / Analyse nodes on the same clip runs windows of
    frames through the engine's clip pipeline (source frames in, denoised frames out); its output must not change
    (radius 1..6, 8 / 16 bit, gray, 4:2:2 / 4:4:0, non-default sharp / rfilter, plane=0 / 4, scene changes, clip ends)"""
    import minivs
    monkeypatch.setenv("MVB200_LOOKAHEAD", str(window))
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    N = len(frames)
    core = minivs.Core()
    mv = core.load(mock_plugin)
    clip = core.source(frames, bits=ck.get("bits", 8), ssw=ck.get("ssw", 1), ssh=ck.get("ssh", 1))
    sup = mv.Super(clip, **skw)
    vn = [mv.Analyse(sup, isb=isb, delta=d, **pipeline._fix(akw)) for d in range(1, radius + 1) for isb in (1, 0)]
    dn = mv.invoke("Degrain%d" % radius, clip, sup, *vn, **dkw)
    dn.set_cache(0)
    before = (sup.computed(), [v.computed() for v in vn])  # what the filters' creation read (frame 0 of every input)
    order = list(range(N)) + [N - 1, 0, N // 2]
    for n in order:
        assert [pipeline.sha(p) for p in dn.get_frame(n).planes()] == GOLD[case]["degrain"][n], "frame %d" % n
    # the Super and Analyse nodes are not asked for a single frame on this path
    assert (sup.computed(), [v.computed() for v in vn]) == before


@pytest.mark.parametrize("window", [2, 7])
@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3", "p10_444_blk8_pel4_bicubic", "noise_badblocks_blk8", "p8_420_blk16_ov4_uncovered"])
def test_flowfps_fused_path_glue(mock_plugin, monkeypatch, case, window):
    """MVB200_LOOKAHEAD: mv.FlowFPS fed by this plugin's own Super / Analyse nodes produces windows of OUTPUT frames with one
    call of the engine's FlowFPS clip pipeline; every variant (mask 0 / 1 / 2, rates, blend on / off, scene changes)
    must give the reference's frames, in order and out of order"""
    import minivs
    monkeypatch.setenv("MVB200_LOOKAHEAD", str(window))
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    nplanes = 1 if ck.get("gray") else 3
    core = minivs.Core()
    mv = core.load(mock_plugin)
    clip = core.source(frames, bits=ck.get("bits", 8), ssw=ck.get("ssw", 1), ssh=ck.get("ssh", 1))
    sup = mv.Super(clip, **skw)
    vn = [mv.Analyse(sup, isb=isb, delta=1, **pipeline._fix(akw)) for isb in (1, 0)]
    nodes = [mv.FlowFPS(clip, sup, vn[0], vn[1], **{k: (float(x) if k == "ml" else x) for k, x in kw.items()}) for kw in pipeline_flow.FLOW[case]["fps"]]
    before = (sup.computed(), [v.computed() for v in vn])  # what the filters' creation read (frame 0 of every input)
    # the node-by-node evaluation of the same script, for the frame properties (Super_* on frames built on super frame 0)
    monkeypatch.setenv("MVB200_LOOKAHEAD", "0")
    core2 = minivs.Core()
    mv2 = core2.load(mock_plugin)
    clip2 = core2.source(frames, bits=ck.get("bits", 8), ssw=ck.get("ssw", 1), ssh=ck.get("ssh", 1))
    sup2 = mv2.Super(clip2, **skw)
    vn2 = [mv2.Analyse(sup2, isb=isb, delta=1, **pipeline._fix(akw)) for isb in (1, 0)]
    plain = [mv2.FlowFPS(clip2, sup2, vn2[0], vn2[1], **{k: (float(x) if k == "ml" else x) for k, x in kw.items()}) for kw in pipeline_flow.FLOW[case]["fps"]]
    for i, node in enumerate(nodes):
        node.set_cache(0)
        M = node.num_frames
        for n in list(range(M)) + [M - 1, 0, M // 2]:
            fr = node.get_frame(n)
            assert [pipeline.sha(p) for p in fr.planes()[:nplanes]] == GOLD_FLOW[case]["fps"][i][n], "variant %d frame %d" % (i, n)
            if n < 3:
                rf = plain[i].get_frame(n)
                for key in ("Super_pel", "Super_levels", "_DurationNum", "_DurationDen"):
                    assert fr.prop_int(key) == rf.prop_int(key), "variant %d frame %d: property %s" % (i, n, key)
    assert (sup.computed(), [v.computed() for v in vn]) == before


@pytest.mark.parametrize("window", [2, 5])
@pytest.mark.parametrize("case", sorted(pipeline.EXTRA))
def test_compensate_fused_path_glue(mock_plugin, monkeypatch, case, window):
    """MVB200_LOOKAHEAD: mv.Compensate on this plugin's own Super / Analyse nodes produces windows of frames with one call of
    the engine's Compensate clip pipeline; the goldens' variants (time, thsad, scbehavior, scene-change thresholds)"""
    import minivs
    monkeypatch.setenv("MVB200_LOOKAHEAD", str(window))
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    N = len(frames)
    nplanes = 1 if ck.get("gray") else 3
    core = minivs.Core()
    mv = core.load(mock_plugin)
    clip = core.source(frames, bits=ck.get("bits", 8), ssw=ck.get("ssw", 1), ssh=ck.get("ssh", 1))
    sup = mv.Super(clip, **skw)
    vec = mv.Analyse(sup, isb=1, delta=1, **pipeline._fix(akw))
    cn = mv.Compensate(clip, sup, vec, **pipeline.EXTRA[case]["comp"])
    cn.set_cache(0)
    before = (sup.computed(), vec.computed())
    for n in list(range(N)) + [N - 1, 0, N // 2]:
        assert [pipeline.sha(p) for p in cn.get_frame(n).planes()[:nplanes]] == GOLD[case]["compensate"][n], "frame %d" % n
    assert (sup.computed(), vec.computed()) == before
    # frame properties: the node-by-node path builds the output on super frame n, so frame 0 carries Super_* (MVCompensate.c:113)
    monkeypatch.setenv("MVB200_LOOKAHEAD", "0")
    core2 = minivs.Core()
    mv2 = core2.load(mock_plugin)
    clip2 = core2.source(frames, bits=ck.get("bits", 8), ssw=ck.get("ssw", 1), ssh=ck.get("ssh", 1))
    sup2 = mv2.Super(clip2, **skw)
    plain = mv2.Compensate(clip2, sup2, mv2.Analyse(sup2, isb=1, delta=1, **pipeline._fix(akw)), **pipeline.EXTRA[case]["comp"])
    for n in (0, 1):
        for key in ("Super_pel", "Super_levels", "Super_hpad", "Super_modeyuv"):
            assert cn.get_frame(n).prop_int(key) == plain.get_frame(n).prop_int(key), "frame %d: property %s" % (n, key)


@pytest.mark.parametrize("window", [2, 5])
@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3", "p10_444_blk8_pel4_bicubic", "gray_blk4_pel1_rf0", "noise_badblocks_blk8", "p8_420_blk16_ov4_uncovered"])
def test_flow_consumers_fused_path_glue(mock_plugin, monkeypatch, case, window):
    """MVB200_LOOKAHEAD: mv.Flow / mv.FlowInter / mv.FlowBlur fed by this plugin's own Super / Analyse nodes produce windows
    of frames with one call of the engine's clip pipelines (mvb200_flow_clip_host, _flowinter_, _flowblur_): the
    reference's frames for every variant of the goldens (modes, times, scene changes, clip ends), in and out of order,
    and the frame properties of the node-by-node evaluation (a compensated frame carries the REFERENCE frame's
    properties, plus Super_* when that is frame 0)"""
    import minivs
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    N = len(frames)
    nplanes = 1 if ck.get("gray") else 3
    tags = [{"Tag": 100 + n} for n in range(N)]
    F = pipeline_flow.FLOW[case]

    def build(fused):
        if fused:
            monkeypatch.setenv("MVB200_LOOKAHEAD", str(window))
        else:
            monkeypatch.setenv("MVB200_LOOKAHEAD", "0")
        core = minivs.Core()
        mv = core.load(mock_plugin)
        clip = core.source(frames, bits=ck.get("bits", 8), ssw=ck.get("ssw", 1), ssh=ck.get("ssh", 1), props=tags)
        sup = mv.Super(clip, **skw)
        vn = [mv.Analyse(sup, isb=isb, delta=1, **pipeline._fix(akw)) for isb in (1, 0)]
        nodes = []
        for kw in F.get("flow", []):
            v, kw = pipeline_flow._split(kw)
            nodes.append(("flow", len([x for x in nodes if x[0] == "flow"]), bool(kw.get("fields")), mv.Flow(clip, sup, vn[v], **{k: (float(x) if k == "time" else x) for k, x in kw.items()})))
        for i, kw in enumerate(F.get("inter", [])):
            nodes.append(("inter", i, False, mv.FlowInter(clip, sup, vn[0], vn[1], **{k: (float(x) if k in ("time", "ml") else x) for k, x in kw.items()})))
        for i, kw in enumerate(F.get("blur", [])):
            nodes.append(("blur", i, False, mv.FlowBlur(clip, sup, vn[0], vn[1], **{k: (float(x) if k == "blur" else x) for k, x in kw.items()})))
        return core, sup, vn, nodes

    plain = build(False)
    fused = build(True)
    before = (fused[1].computed(), [v.computed() for v in fused[2]])  # what the filters' creation read (frame 0 of every input)
    for (stage, i, fields, node), (_, _, _, ref) in zip(fused[3], plain[3]):
        node.set_cache(0)
        for n in list(range(N)) + [N - 1, 0, N // 2]:
            fr, rf = node.get_frame(n), ref.get_frame(n)
            assert [pipeline.sha(p) for p in fr.planes()[:nplanes]] == GOLD_FLOW[case][stage][i][n], "%s variant %d frame %d" % (stage, i, n)
            for key in ("Tag", "Super_pel", "Super_height", "Super_levels"):
                assert fr.prop_int(key) == rf.prop_int(key), "%s variant %d frame %d: property %s" % (stage, i, n, key)
    # on the fused paths the Super and Analyse nodes are not asked for a single frame (mv.Flow with fields keeps the nodes)
    if not any(fields for _, _, fields, _ in fused[3]):
        assert (fused[1].computed(), [v.computed() for v in fused[2]]) == before


@pytest.mark.parametrize("window", [2, 7])
@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3", "gray_blk4_pel1_rf0", "noise_badblocks_blk8", "p8_420_blk16_ov4_uncovered", "blk64_420_ov32_r2"])
def test_blockfps_fused_path_glue(mock_plugin, monkeypatch, case, window):
    """MVB200_LOOKAHEAD: mv.BlockFPS fed by this plugin's own Super / Analyse nodes produces windows of OUTPUT frames with one
    call of the engine's BlockFPS clip pipeline (mvb200_blockfps_clip_host): every variant of the goldens (modes 0..8,
    rates, blend on / off, scene changes), in and out of order, with the node-by-node evaluation's frame properties"""
    import minivs
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    nplanes = 1 if ck.get("gray") else 3
    tags = [{"Tag": 100 + n} for n in range(len(frames))]

    def build(fused):
        if fused:
            monkeypatch.setenv("MVB200_LOOKAHEAD", str(window))
        else:
            monkeypatch.setenv("MVB200_LOOKAHEAD", "0")
        core = minivs.Core()
        mv = core.load(mock_plugin)
        clip = core.source(frames, bits=ck.get("bits", 8), ssw=ck.get("ssw", 1), ssh=ck.get("ssh", 1), props=tags)
        sup = mv.Super(clip, **skw)
        vn = [mv.Analyse(sup, isb=isb, delta=1, **pipeline._fix(akw)) for isb in (1, 0)]
        nodes = [mv.BlockFPS(clip, sup, vn[0], vn[1], **{k: (float(x) if k == "ml" else x) for k, x in kw.items()}) for kw in pipeline_flow.FLOW[case]["bfps"]]
        return core, sup, vn, nodes

    plain = build(False)
    fused = build(True)
    before = (fused[1].computed(), [v.computed() for v in fused[2]])  # what the filters' creation read (frame 0 of every input)
    for i, (node, ref) in enumerate(zip(fused[3], plain[3])):
        node.set_cache(0)
        M = node.num_frames
        for n in list(range(M)) + [M - 1, 0, M // 2]:
            fr, rf = node.get_frame(n), ref.get_frame(n)
            assert [pipeline.sha(p) for p in fr.planes()[:nplanes]] == GOLD_FLOW[case]["bfps"][i][n], "variant %d frame %d" % (i, n)
            for key in ("Tag", "Super_pel", "Super_levels", "_DurationNum", "_DurationDen"):
                assert fr.prop_int(key) == rf.prop_int(key), "variant %d frame %d: property %s" % (i, n, key)
    assert (fused[1].computed(), [v.computed() for v in fused[2]]) == before


@pytest.mark.parametrize("window,threads", [(3, 4), (2, 8)])
def test_lookahead_paths_under_concurrent_requests(mock_plugin, monkeypatch, window, threads):
    """fmParallel: several host threads pull frames at once (mini host's render); the window caches of mv.Analyse and of
    the fused DegrainN must neither deadlock nor serve a wrong frame afterwards.  With more threads than the cache has
    windows, a frame that was cached when a request started is evicted before the request finishes: the request keeps it."""
    import minivs
    monkeypatch.setenv("MVB200_LOOKAHEAD", str(window))
    case = "blk16_ov8_pel2_p8_r3"
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    N = len(frames)
    sp, aps, dp = pipeline.params_for(case)
    for rounds in range(2):
        core = minivs.Core()
        mv = core.load(mock_plugin)
        clip = core.source(frames, bits=8)
        sup = mv.Super(clip, **skw)
        vn = [mv.Analyse(sup, isb=isb, delta=d, **akw) for d in range(1, radius + 1) for isb in (1, 0)]
        dn = mv.invoke("Degrain%d" % radius, clip, sup, *vn, **dkw)
        comp = mv.Compensate(clip, sup, vn[0])  # a consumer that goes through the Analyse node (look-ahead cache)
        for node in (dn, comp) + tuple(vn):
            node.set_cache(0)
        dn.render(0, N, threads)
        comp.render(0, N, threads)
        import numpy as np
        for n in range(N):
            assert [pipeline.sha(p) for p in dn.get_frame(n).planes()] == GOLD[case]["degrain"][n]
            blob = np.frombuffer(vn[0].get_frame(n).prop_data("MVTools_vectors"), dtype=np.uint8)
            assert pipeline.sha(pipeline.mask_blob(blob, aps[0])) == GOLD[case]["vectors"][n][0]


@pytest.mark.parametrize("window,threads", [(3, 1), (3, 4), (4, 2)])
def test_fused_windows_are_aligned_pipelined_and_computed_once(mock_plugin, monkeypatch, window, threads):
    """The fused DegrainN cuts the clip into aligned windows; frames pulled in order by any number of threads cost exactly one
    engine call per window (the request that first touches window k also enqueues window k + 1 -- nobody recomputes it), and
    the automatic window (MVB200_LOOKAHEAD unset) gives the same frames."""
    import ctypes
    import minivs
    case = "blk16_ov8_pel2_p8_r3"
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    N = len(frames)
    calls = ctypes.CDLL(mock_plugin).mvb200_mock_clip_calls
    calls.restype = ctypes.c_long
    for la in (str(window), None):
        if la is None:
            monkeypatch.delenv("MVB200_LOOKAHEAD", raising=False)
        else:
            monkeypatch.setenv("MVB200_LOOKAHEAD", la)
        core = minivs.Core()
        mv = core.load(mock_plugin)
        clip = core.source(frames, bits=8)
        sup = mv.Super(clip, **skw)
        vn = [mv.Analyse(sup, isb=isb, delta=d, **akw) for d in range(1, radius + 1) for isb in (1, 0)]
        dn = mv.invoke("Degrain%d" % radius, clip, sup, *vn, **dkw)
        dn.set_cache(0)
        before = calls()
        dn.render(0, N, threads)
        made = calls() - before
        B = window if la else 16  # (the stand-in's plan: the automatic window is the 16-frame minimum)
        assert made == (N + B - 1) // B, "window %s: %d engine calls for %d frames" % (la, made, N)
        for n in range(N):
            assert [pipeline.sha(p) for p in dn.get_frame(n).planes()] == GOLD[case]["degrain"][n]


def test_fused_windows_spread_over_every_device(mock_plugin):
    """With several devices (here: the stand-in pretends to have three) the windows of a fused filter go round the engine
    contexts, one per device, and as many windows as there are devices are enqueued ahead; the frames do not change.
    (A child process: the plugin reads the device list once per process.)"""
    import subprocess
    import sys
    code = r'''
import ctypes, json, os, sys
sys.path[:0] = [%r, %r, %r]
import minivs, pipeline
case = "blk16_ov8_pel2_p8_r3"
ck, skw, akw, radius, dkw = pipeline.CASES[case]
frames = pipeline.make_frames(ck)
core = minivs.Core()
mv = core.load(%r)
clip = core.source(frames, bits=8)
sup = mv.Super(clip, **skw)
vn = [mv.Analyse(sup, isb=isb, delta=d, **akw) for d in range(1, radius + 1) for isb in (1, 0)]
dn = mv.invoke("Degrain%%d" %% radius, clip, sup, *vn, **dkw)
dn.set_cache(0)
dn.render(0, len(frames), 4)
L = ctypes.CDLL(%r)
L.mvb200_mock_contexts.restype = ctypes.c_long
L.mvb200_mock_clip_calls.restype = ctypes.c_long
print(json.dumps({"contexts": L.mvb200_mock_contexts(), "calls": L.mvb200_mock_clip_calls(),
                  "frames": [[pipeline.sha(p) for p in dn.get_frame(n).planes()] for n in range(len(frames))]}))
''' % (os.path.join(ROOT, "vapoursynth-mvtools_b200"), os.path.join(ROOT, "host", "minivs"), HERE, mock_plugin, mock_plugin)
    env = dict(os.environ, MVB200_MOCK_DEVICES="3", MVB200_LOOKAHEAD="2", MVB200_CONTEXTS_PER_DEVICE="1")
    env.pop("MVB200_DEVICE", None)
    env.pop("MVB200_DEVICES", None)
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = json.loads(r.stdout.strip().splitlines()[-1])
    case = "blk16_ov8_pel2_p8_r3"
    assert got["frames"] == GOLD[case]["degrain"]
    assert got["contexts"] == 3, "windows were not spread over the three devices"
    assert got["calls"] >= 4


def test_a_failed_window_reports_its_error_and_nothing_else_breaks(mock_plugin, monkeypatch):
    """The engine call of ONE window fails (the stand-in injects it): every frame of that window raises the engine's message through
    the frame server -- no hang, no crash, no frame from a half-written buffer --, the other windows' frames are the golden's, and
    a later request for the failed window's frames fails again rather than serving stale data."""
    import minivs
    monkeypatch.setenv("MVB200_LOOKAHEAD", "3")
    monkeypatch.setenv("MVB200_MOCK_FAIL_CLIP_CALL", str(ctypes_calls(mock_plugin) + 2))  # the second window of this script
    case = "blk16_ov8_pel2_p8_r3"
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    N = len(frames)
    core = minivs.Core()
    mv = core.load(mock_plugin)
    clip = core.source(frames, bits=8)
    sup = mv.Super(clip, **skw)
    vn = [mv.Analyse(sup, isb=isb, delta=d, **akw) for d in range(1, radius + 1) for isb in (1, 0)]
    dn = mv.invoke("Degrain%d" % radius, clip, sup, *vn, **dkw)
    dn.set_cache(0)
    failed, good = [], []
    for n in range(N):
        try:
            fr = dn.get_frame(n)
        except minivs.Error as e:
            assert "injected failure" in str(e) and str(e).startswith("Degrain%d: " % radius), str(e)
            failed.append(n)
            continue
        assert [pipeline.sha(p) for p in fr.planes()] == GOLD[case]["degrain"][n], "frame %d" % n
        good.append(n)
    assert failed == [3, 4, 5], (failed, good)  # exactly the second window of three frames
    with pytest.raises(minivs.Error):
        dn.get_frame(4)


def ctypes_calls(so):
    import ctypes
    f = ctypes.CDLL(so).mvb200_mock_clip_calls
    f.restype = ctypes.c_long
    return f()


def test_fused_windows_on_four_contexts_of_one_device(mock_plugin):
    """The window plan may ask for up to four engine contexts on one device (4K clips: more, smaller windows): four windows are
    enqueued ahead, eight host threads pull frames, windows are evicted while requests hold them -- every frame is still the
    golden's, for a fused DegrainN and for a vector clip read directly.  (A child process: the device list is read once.)"""
    import subprocess
    import sys
    code = r'''
import ctypes, json, os, sys
sys.path[:0] = [%r, %r, %r]
import numpy as np
import minivs, pipeline
case = "blk16_ov8_pel2_p8_r3"
ck, skw, akw, radius, dkw = pipeline.CASES[case]
frames = pipeline.make_frames(ck)
sp, aps, dp = pipeline.params_for(case)
core = minivs.Core()
mv = core.load(%r)
clip = core.source(frames, bits=8)
sup = mv.Super(clip, **skw)
vn = [mv.Analyse(sup, isb=isb, delta=d, **akw) for d in range(1, radius + 1) for isb in (1, 0)]
dn = mv.invoke("Degrain%%d" %% radius, clip, sup, *vn, **dkw)
for node in [dn] + vn:
    node.set_cache(0)
for rounds in range(3):
    dn.render(0, len(frames), 8)
    vn[0].render(0, len(frames), 8)
L = ctypes.CDLL(%r)
L.mvb200_mock_contexts.restype = ctypes.c_long
print(json.dumps({"contexts": L.mvb200_mock_contexts(),
                  "frames": [[pipeline.sha(p) for p in dn.get_frame(n).planes()] for n in range(len(frames))],
                  "vectors": [pipeline.sha(pipeline.mask_blob(np.frombuffer(vn[0].get_frame(n).prop_data("MVTools_vectors"), dtype=np.uint8), aps[0])) for n in range(len(frames))]}))
''' % (os.path.join(ROOT, "vapoursynth-mvtools_b200"), os.path.join(ROOT, "host", "minivs"), HERE, mock_plugin, mock_plugin)
    env = dict(os.environ, MVB200_LOOKAHEAD="2", MVB200_CONTEXTS_PER_DEVICE="4")
    for k in ("MVB200_DEVICE", "MVB200_DEVICES", "MVB200_MOCK_DEVICES"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = json.loads(r.stdout.strip().splitlines()[-1])
    case = "blk16_ov8_pel2_p8_r3"
    assert got["frames"] == GOLD[case]["degrain"]
    assert got["vectors"] == [v[0] for v in GOLD[case]["vectors"]]
    assert got["contexts"] == 4, "the windows did not go round four contexts"


@pytest.mark.parametrize("window", ["3", None])
def test_recalculate_passes_inside_the_fused_chains(mock_plugin, have_ref, monkeypatch, window):
    """The refinemotion shape -- Analyse -> Recalculate (-> Recalculate) -> DegrainN / Compensate -- is still one engine call per
    window: the fused paths recognise this plugin's own mv.Recalculate nodes on top of its mv.Analyse nodes and hand their
    parameters to the clip pipeline; the frames are the reference plugin's, and none of the Super / Analyse / Recalculate nodes
    is asked for a frame."""
    if not have_ref:
        pytest.skip("oracle/_ref/libmvtools_ref.so not built here (no /root/reference)")
    import minivs
    if window:
        monkeypatch.setenv("MVB200_LOOKAHEAD", window)
    else:
        monkeypatch.delenv("MVB200_LOOKAHEAD", raising=False)
    case = "blk16_ov8_pel2_p8_r3"
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    N = len(frames)
    outs, counts = [], None
    for path in (minivs.REF_PLUGIN, mock_plugin):
        core = minivs.Core()
        mv = core.load(path)
        clip = core.source(frames, bits=8)
        sup = mv.Super(clip, **skw)
        vb, vf = mv.Analyse(sup, isb=1, delta=1, **akw), mv.Analyse(sup, isb=0, delta=1, **akw)
        rb = mv.Recalculate(sup, vb, blksize=8, overlap=4, thsad=100)
        rb2 = mv.Recalculate(sup, rb, blksize=8, overlap=4, thsad=50, search=3, searchparam=1)
        rf = mv.Recalculate(sup, vf, blksize=8, overlap=4, thsad=100, divide=0)
        dn = mv.Degrain1(clip, sup, rb2, rf, thsad=300)
        comp = mv.Compensate(clip, sup, rb)
        nodes = (sup, vb, vf, rb, rb2, rf)
        before = [x.computed() for x in nodes]
        res = []
        for n in list(range(N)) + [0, N - 1]:
            res.append(([pipeline.sha(p) for p in dn.get_frame(n).planes()], [pipeline.sha(p) for p in comp.get_frame(n).planes()]))
        outs.append(res)
        if path == mock_plugin:
            counts = (before, [x.computed() for x in nodes])
    assert outs[0] == outs[1]
    assert counts[0] == counts[1], "the fused path asked the Super / Analyse / Recalculate nodes for frames"


@pytest.mark.parametrize("window", ["3", None])
@pytest.mark.parametrize("case", ["c1_640x360_p8_blk8_pel1", "blk16_ov8_pel2_p16_r3", "gray_blk4_pel1_rf0", "divide2_blk16_esa_neg_badrange", "super_nochroma_blk8_r2"])
def test_vector_clip_windows_from_source_frames(mock_plugin, monkeypatch, case, window):
    """A vector clip that is read directly (the script's output, mv.Mask, mv.SCDetection, another plugin): when its super clip is
    this plugin's own mv.Super, windows of vector frames come from ONE engine call over SOURCE frames (mvb200_analyse_clip_host) --
    pixels (the super frame), MVTools_vectors, MVTools_MVAnalysisData and Super_* identical to the goldens / the node-by-node
    evaluation, and the Super node is asked for frame 0 only (the single chain that serves the script-building reads)."""
    import minivs
    import numpy as np
    if window:
        monkeypatch.setenv("MVB200_LOOKAHEAD", window)
    else:
        monkeypatch.delenv("MVB200_LOOKAHEAD", raising=False)
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    N = len(frames)
    sp, aps, dp = pipeline.params_for(case)
    core = minivs.Core()
    mv = core.load(mock_plugin)
    clip = core.source(frames, bits=ck.get("bits", 8), ssw=ck.get("ssw", 1), ssh=ck.get("ssh", 1))
    sup = mv.Super(clip, **skw)
    vec = mv.Analyse(sup, isb=1, delta=1, **pipeline._fix(akw))
    scd = mv.SCDetection(clip, vec)  # a consumer that reads the vector clip frame by frame
    vec.set_cache(0)
    for n in list(range(N)) + [N - 1, 0, N // 2]:
        fr = vec.get_frame(n)
        assert [pipeline.sha(p) for p in fr.planes()] == GOLD[case]["super"][n], "frame %d: pixels" % n
        blob = np.frombuffer(fr.prop_data("MVTools_vectors"), dtype=np.uint8)
        assert pipeline.sha(pipeline.mask_blob(blob, aps[0])) == GOLD[case]["vectors"][n][0], "frame %d: vectors" % n
        assert fr.prop_data("MVTools_MVAnalysisData") is not None
        assert (fr.prop_int("Super_pel") is not None) == (n == 0), "frame %d: Super_* properties" % n
        scd.get_frame(n)
    assert sup.computed() <= 2, "the Super node was asked for more than the frames of the single chain of frame 0"


def test_window_buffers_are_recycled_in_steady_state(mock_plugin):
    """A vector clip's windows have a small source buffer and a large output buffer (super frames): after the first windows the
    pinned buffers go round -- no allocation per window (page-locking a gigabyte costs half a second on the GPU box).  Own
    process: the pool is process-wide and the stand-in's compute is switched off."""
    code = """
import ctypes, os, sys
sys.path[:0] = [%r, %r, %r]
import minivs, pipeline
so = %r
core = minivs.Core()
mv = core.load(so)
allocs = ctypes.CDLL(so).mvb200_mock_pinned_allocs
ck = dict(width=160, height=96, nframes=8, bits=8)
frames = pipeline.make_frames(ck) * 40
clip = core.source(frames, bits=8, ssw=1, ssh=1)
vec = mv.Analyse(mv.Super(clip, pel=2, hpad=8, vpad=8), isb=1, blksize=8, overlap=4)
vec.set_cache(0)
vec.render(0, 96, 4)
a0 = allocs()
vec.render(96, 304, 4)
a1 = allocs()
print("ALLOCS", a0, a1)
assert a0 > 0 and a1 - a0 <= 4, (a0, a1)  # 13 more windows: a late extra window in flight may add a pair, a per-window allocation adds 26
""" % (os.path.join(ROOT, "host", "minivs"), HERE, os.path.join(ROOT, "vapoursynth-mvtools_b200"), mock_plugin)
    env = dict(os.environ, MVB200_LOOKAHEAD="16", MVB200_MOCK_NOCOMPUTE="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_non_canonical_chains_keep_the_node_by_node_evaluation(mock_plugin, monkeypatch):
    """MVB200_LOOKAHEAD set, but the script is not the canonical chain: vectors searched on ANOTHER Super node of the same
    clip, a vector clip with divide feeding mv.Flow, a consumer with fields -- the fused paths must decline (the Super /
    Analyse nodes are asked for frames) and the results stay the reference's"""
    import minivs
    monkeypatch.setenv("MVB200_LOOKAHEAD", "4")
    # (a) another Super instance
    case = "blk8_ov4_pel2_p8"
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    N = len(frames)
    core = minivs.Core()
    mv = core.load(mock_plugin)
    clip = core.source(frames, bits=8)
    sup, sup2 = mv.Super(clip, **skw), mv.Super(clip, **skw)
    vn = [mv.Analyse(sup2, isb=isb, delta=1, **akw) for isb in (1, 0)]
    dn = mv.Degrain1(clip, sup, vn[0], vn[1])
    inter = mv.FlowInter(clip, sup, vn[0], vn[1])
    before = sup.computed()
    for n in range(N):
        assert [pipeline.sha(p) for p in dn.get_frame(n).planes()] == GOLD[case]["degrain"][n]
        assert [pipeline.sha(p) for p in inter.get_frame(n).planes()] == GOLD_FLOW[case]["inter"][0][n]
    assert sup.computed() > before
    # a consumer with fields on the canonical nodes
    sup3 = mv.Super(clip, **skw)
    v3 = mv.Analyse(sup3, isb=1, delta=1, **akw)
    kw = dict(pipeline_flow.FLOW[case]["flow"][4])  # fields=1, tff=1
    assert kw.get("fields") == 1 and "vec" not in kw
    fl = mv.Flow(clip, sup3, v3, **{k: (float(x) if k == "time" else x) for k, x in kw.items()})
    before = v3.computed()
    for n in range(N):
        assert [pipeline.sha(p) for p in fl.get_frame(n).planes()] == GOLD_FLOW[case]["flow"][4][n]
    assert v3.computed() > before
    # (b) divide: mv.Flow declines, mv.Compensate (a block filter) does not
    case = "divide2_blk16_esa_neg_badrange"
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    core = minivs.Core()
    mv = core.load(mock_plugin)
    clip = core.source(frames, bits=8)
    sup = mv.Super(clip, **skw)
    vec = mv.Analyse(sup, isb=1, delta=1, **pipeline._fix(akw))
    kw = dict(pipeline_flow.FLOW[case]["flow"][0])
    fl = mv.Flow(clip, sup, vec, **kw)
    before = vec.computed()  # (the vector clip: creating the filters already made its look-ahead read the super frames)
    for n in range(len(frames)):
        assert [pipeline.sha(p) for p in fl.get_frame(n).planes()] == GOLD_FLOW[case]["flow"][0][n]
    assert vec.computed() > before
    comp = mv.Compensate(clip, sup, vec, **pipeline.EXTRA[case]["comp"])
    before = vec.computed()
    for n in range(len(frames)):
        assert [pipeline.sha(p) for p in comp.get_frame(n).planes()] == GOLD[case]["compensate"][n]
    assert vec.computed() == before


def test_fused_request_survives_the_eviction_of_its_frame(mock_plugin, monkeypatch):
    """A request that finds its frame in the window cache keeps it: while that request waits for the (here: slow) source
    frame it takes its properties from, other requests' windows push the frame out of the cache (two windows are kept)."""
    import threading
    import time
    import clipgen
    import minivs
    monkeypatch.setenv("MVB200_LOOKAHEAD", "2")
    frames = clipgen.make_clip(96, 64, 10, bits=8)
    core = minivs.Core()
    mv = core.load(mock_plugin)
    clip = core.source(frames, bits=8, delay_ms={1: 1500})
    sup = mv.Super(clip, pel=1)
    vn = [mv.Analyse(sup, isb=isb, delta=1, blksize=8) for isb in (1, 0)]
    dn = mv.Degrain1(clip, sup, vn[0], vn[1])
    dn.set_cache(0)
    first = [p.copy() for p in dn.get_frame(1).planes()]  # the window [1, 3) is computed and cached
    result = {}

    def late():
        try:
            result["planes"] = dn.get_frame(1).planes()  # a cache hit that then waits 1.5 s for source frame 1
        except Exception as e:  # noqa: BLE001
            result["error"] = str(e)

    t = threading.Thread(target=late)
    t.start()
    time.sleep(0.3)
    dn.get_frame(4)  # windows [4, 6) and [6, 8): four newer frames, [1, 3) is evicted (neither needs source frame 1)
    dn.get_frame(6)
    t.join()
    assert "error" not in result, result.get("error")
    assert all((a == b).all() for a, b in zip(result["planes"], first))


def test_analyse_lookahead_batches_the_engine_calls(mock_plugin, monkeypatch):
    """with a window of 4, six frames requested in order cost two window searches per vector clip, not six"""
    import minivs
    monkeypatch.setenv("MVB200_LOOKAHEAD", "4")
    case = "blk8_ov4_pel2_p8"
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    core = minivs.Core()
    mv = core.load(mock_plugin)
    clip = core.source(frames, bits=8)
    sup = mv.Super(clip, **skw)
    vec = mv.Analyse(sup, isb=1, delta=1, **akw)
    blobs = [vec.get_frame(n).prop_data("MVTools_vectors") for n in range(len(frames))]
    sp, aps, dp = pipeline.params_for(case)
    import numpy as np
    assert [pipeline.sha(pipeline.mask_blob(np.frombuffer(b, dtype=np.uint8), aps[0])) for b in blobs] == [row[0] for row in GOLD[case]["vectors"]]
    # out-of-order and repeated requests are served as well (cache hits, evictions, single-frame fall-back)
    for n in (5, 0, 3, 3, 1):
        assert vec.get_frame(n).prop_data("MVTools_vectors") == blobs[n]


def test_field_property_glue_matches_reference_live(mock_plugin, have_ref):
    """field-separated clips: without a tff argument the filters read the _Field property of the frames
    (MVAnalyse.c:136-179, MVCompensate.c, MVFlow.cpp:268-299); the plugin's glue against the compiled reference"""
    if not have_ref:
        pytest.skip("oracle/_ref/libmvtools_ref.so not built here (no /root/reference)")
    import minivs
    case = "blk8_ov4_pel2_p8"
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    N = len(frames)
    props = [{"_Field": (n + 1) % 2} for n in range(N)]
    outs = []
    for path in (minivs.REF_PLUGIN, mock_plugin):
        core = minivs.Core()
        mv = core.load(path)
        clip = core.source(frames, bits=8, props=props)
        sup = mv.Super(clip, **skw)
        vb = mv.Analyse(sup, isb=1, delta=1, fields=1, **akw)
        vf = mv.Analyse(sup, isb=0, delta=1, fields=1, **akw)
        comp = mv.Compensate(clip, sup, vb, fields=1)
        flow = mv.Flow(clip, sup, vf, fields=1, time=60.0)
        rec = mv.Recalculate(sup, vb, blksize=8, overlap=4, fields=1, thsad=50)
        res = []
        for n in range(N):
            res.append((vb.get_frame(n).prop_data("MVTools_vectors"), vf.get_frame(n).prop_data("MVTools_vectors"), rec.get_frame(n).prop_data("MVTools_vectors"),
                        [pipeline.sha(p) for p in comp.get_frame(n).planes()], [pipeline.sha(p) for p in flow.get_frame(n).planes()]))
        outs.append(res)
    assert outs[0] == outs[1]
    # and the error when neither the property nor tff is there
    core = minivs.Core()
    mv = core.load(mock_plugin)
    sup = mv.Super(core.source(frames, bits=8), **skw)
    with pytest.raises(minivs.Error) as e:
        mv.Analyse(sup, isb=1, fields=1, **akw).get_frame(1)
    assert "_Field property not found in input frame" in str(e.value)


@pytest.mark.parametrize("lookahead", [None, "3"])
def test_static_reference_frame_glue_matches_reference_live(mock_plugin, have_ref, monkeypatch, lookahead):
    """delta < 0: every frame is searched against ONE fixed frame (-delta; MVAnalyse.c:84-109) -- the vector clip, and
    Compensate / Flow / Mask / SCDetection / Recalculate on it, against the compiled reference; also with the look-ahead
    set (windows of mv.Analyse share the reference frame; the fused paths decline absolute references)"""
    if not have_ref:
        pytest.skip("oracle/_ref/libmvtools_ref.so not built here (no /root/reference)")
    import minivs
    if lookahead:
        monkeypatch.setenv("MVB200_LOOKAHEAD", lookahead)
    else:
        monkeypatch.delenv("MVB200_LOOKAHEAD", raising=False)
    case = "blk8_ov4_pel2_p8"
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    N = len(frames)
    outs = []
    for path in (minivs.REF_PLUGIN, mock_plugin):
        core = minivs.Core()
        mv = core.load(path)
        clip = core.source(frames, bits=8)
        sup = mv.Super(clip, **skw)
        vec = mv.Analyse(sup, isb=0, delta=-2, **akw)
        comp = mv.Compensate(clip, sup, vec)
        flow = mv.Flow(clip, sup, vec, time=70.0)
        mask = mv.Mask(clip, vec, kind=1)
        scd = mv.SCDetection(clip, vec)
        rec = mv.Recalculate(sup, vec, blksize=8, overlap=4, thsad=50)
        res = []
        for n in range(N):
            res.append((vec.get_frame(n).prop_data("MVTools_vectors"), rec.get_frame(n).prop_data("MVTools_vectors"), [pipeline.sha(p) for p in comp.get_frame(n).planes()],
                        [pipeline.sha(p) for p in flow.get_frame(n).planes()], [pipeline.sha(p) for p in mask.get_frame(n).planes()], scd.get_frame(n).prop_int("_SceneChangePrev")))
        outs.append(res)
    for n in range(N):
        for k, what in enumerate(("vectors", "recalculated vectors", "Compensate", "Flow", "Mask", "SCDetection")):
            assert outs[0][n][k] == outs[1][n][k], "frame %d: %s" % (n, what)

