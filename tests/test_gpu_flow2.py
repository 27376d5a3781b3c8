"""GPU parity tests of mv.FlowBlur, mv.BlockFPS, mvb200_analyse_frames, the plugin's look-ahead (MVB200_LOOKAHEAD) and
the fused *_clip_host entry points (FlowFPS, Compensate): CUDA output against the reference's goldens, per frame, through
the plugin script and at full size.  All 36 passed on a B200 (profiles/r1_pytest_gpu_pending.log -- the file was called
test_zz_gpu_pending.py while they were waiting for a GPU)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import mvtools_b200 as mb
import pipeline
import pipeline_flow
from test_gpu_parity import _device_clip, _planes

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_flow.json")))
GOLD_MAIN = json.load(open(os.path.join(HERE, "golden", "golden.json")))
GOLD_FLOW = GOLD


@pytest.mark.parametrize("case", sorted(c for c in pipeline_flow.FLOW if pipeline_flow.FLOW[c].get("blur") and c not in pipeline.LATE))
def test_cuda_flowblur_matches_reference_goldens(gpu_ctx, case):
    """mv.FlowBlur: per-frame entry point and the device-resident batch entry point"""
    import torch
    L = mb.lib()
    ctx = gpu_ctx
    assert pipeline_flow.run_engine_blur(case, ctx.flowblur_frame) == GOLD[case]["blur"]
    frames, sups, vecs = pipeline_flow.inputs(case)
    N = len(frames)
    sp, aps, P = pipeline_flow.flow_params_for(case)
    cl, sl = mb.clip_layout(sp), mb.super_layout(sp)
    dsrc, dsup = _device_clip(frames, cl), _device_clip(sups, sl)
    vs = (len(vecs[0][0]) + 255) // 256 * 256
    dvec = []
    for v in range(2):
        host = np.zeros((N, vs), dtype=np.uint8)
        for n in range(N):
            host[n, : len(vecs[v][n])] = vecs[v][n]
        dvec.append(torch.from_numpy(host).cuda())
    for i, p in enumerate(P["blur"]):
        ddst = torch.full((N, cl.frame_bytes), 0xA5, dtype=torch.uint8, device="cuda")
        mb._check(L.mvb200_flowblur_batch(ctx.handle, C.byref(p), C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), 0, N, dvec[0].data_ptr(), dvec[1].data_ptr(), vs, ddst.data_ptr()))
        ctx.synchronize()
        out = ddst.cpu().numpy()
        for n in range(N):
            assert [pipeline.sha(q) for q in _planes(out[n], cl)] == GOLD[case]["blur"][i][n], "blur variant %d frame %d" % (i, n)

@pytest.mark.parametrize("case", sorted(c for c in pipeline_flow.FLOW if pipeline_flow.FLOW[c].get("bfps") and c not in pipeline.LATE))
def test_cuda_blockfps_matches_reference_goldens(gpu_ctx, case):
    """mv.BlockFPS: every output frame of every variant (modes 0..8, with and without overlap)"""
    ctx = gpu_ctx
    assert pipeline_flow.run_engine_bfps(case, ctx.blockfps_frame, ctx.blend_frame) == GOLD[case]["bfps"]


def test_interpolation_full_size_1080p_properties(gpu_ctx):
    """1920x1080 YUV420P8, pel 2, blk 16 / overlap 8: with a zero vector field between two identical frames every
    interpolation filter must return that frame, whatever its time / mode (all of them are convex blends of
    motion-compensated samples); a full-size run of mv.FlowInter / mv.FlowBlur / mv.BlockFPS and their masks."""
    ctx = gpu_ctx
    W, H, N = 1920, 1080, 3
    import clipgen
    frames = clipgen.make_clip(W, H, 1, bits=8, scene_cut=False)
    sp = mb.super_params(W, H, 8, 1, 1, pel=2)
    sup = ctx.super_frame(sp, frames[0])
    aps = [mb.analyse_params(sp, N, isb=isb, blksize=16, overlap=8) for isb in (1, 0)]
    v0 = [ctx.analyse_frame(ap, sp, sup, sup) for ap in aps]
    th = dict(thscd1=16000, thscd2=255)
    ip = mb.flowinter_params(sp, aps[0].ad, aps[1].ad, N, time=37.0, **th)
    out = ctx.flowinter_frame(ip, sp, ip.time256, sup, sup, v0[0], v0[1], v0[0], v0[1])
    assert out is not None and all(np.array_equal(a, b) for a, b in zip(out, frames[0])), "FlowInter"
    bp = mb.flowblur_params(sp, aps[0].ad, aps[1].ad, N, blur=100.0, **th)
    out = ctx.flowblur_frame(bp, sp, sup, v0[0], v0[1])
    assert out is not None and all(np.array_equal(a, b) for a, b in zip(out, frames[0])), "FlowBlur"
    for mode in (0, 1, 2, 3, 4, 6, 7):
        q = mb.blockfps_params(sp, aps[0].ad, aps[1].ad, N, W, H, num=60, den=1, mode=mode, **th)
        out = ctx.blockfps_frame(q, sp, 100, sup, sup, v0[0], v0[1])
        assert out is not None and all(np.array_equal(a, b) for a, b in zip(out, frames[0])), "BlockFPS mode %d" % mode
    # the blend the filters fall back to at scene changes: time256 = 0 / 256 give the two frames themselves
    other = [255 - p for p in frames[0]]
    assert all(np.array_equal(a, b) for a, b in zip(ctx.blend_frame(sp, 0, frames[0], other), frames[0]))
    assert all(np.array_equal(a, b) for a, b in zip(ctx.blend_frame(sp, 256, frames[0], other), other))

@pytest.mark.parametrize("case", ["blk16_ov8_pel2_p8_r3", "blk16_ov8_pel2_p16_r3", "noise_badblocks_blk8", "divide2_blk16_esa_neg_badrange"])
def test_analyse_frames_matches_goldens(gpu_ctx, case):
    """mvb200_analyse_frames (the plugin's look-ahead: all frames of a vector clip searched in ONE engine call, the
    super frames uploaded once) gives the reference's blobs"""
    ctx = gpu_ctx
    frames = pipeline.make_frames(pipeline.CASES[case][0])
    N = len(frames)
    sp, aps, dp = pipeline.params_for(case)
    sups = [ctx.super_frame(sp, f) for f in frames]
    for i, ap in enumerate(aps):
        jobs = [(n, pipeline.nref_of(ap, n, N), 0) for n in range(N)]
        blobs = ctx.analyse_frames(ap, sp, sups, jobs)
        assert [pipeline.sha(pipeline.mask_blob(b, ap)) for b in blobs] == [row[i] for row in GOLD_MAIN[case]["vectors"]], "vector clip %d" % i
    # a sub-window with its own frame list (indices are positions in the list, not frame numbers)
    ap = aps[0]
    blobs = ctx.analyse_frames(ap, sp, [sups[2], sups[3], sups[1]], [(0, 1, 0), (2, 0, 0), (1, None, 0)])
    assert np.array_equal(blobs[0], ctx.analyse_frame(ap, sp, sups[2], sups[3]))
    assert np.array_equal(blobs[1], ctx.analyse_frame(ap, sp, sups[1], sups[2]))
    assert np.array_equal(blobs[2], ctx.analyse_frame(ap, sp, sups[3], None))

@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3", "noise_badblocks_blk8"])
def test_flowblur_script_through_plugin_matches_reference(built, case):
    import pipeline_flow
    got = pipeline_flow.run_reference_flow(case, plugin_path=mb.PLUGIN_PATH, only=("blur",))
    assert got["blur"] == GOLD_FLOW[case]["blur"], "%s: plugin mv.FlowBlur differs from the reference plugin" % case

@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3", "c1_640x360_p8_blk8_pel1", "noise_badblocks_blk8"])
def test_blockfps_script_through_plugin_matches_reference(built, case):
    import pipeline_flow
    got = pipeline_flow.run_reference_flow(case, plugin_path=mb.PLUGIN_PATH, only=("bfps",))
    assert got["bfps"] == GOLD_FLOW[case]["bfps"], "%s: plugin mv.BlockFPS differs from the reference plugin" % case

@pytest.mark.parametrize("case", ["blk16_ov8_pel2_p8_r3", "divide2_blk16_esa_neg_badrange"])
def test_script_through_plugin_with_lookahead(built, monkeypatch, case):
    """MVB200_LOOKAHEAD=4: mv.Analyse searches four frames per engine call; the script's results do not change"""
    monkeypatch.setenv("MVB200_LOOKAHEAD", "4")
    got = pipeline.run_reference(case, plugin_path=mb.PLUGIN_PATH)
    for stage in ("super", "vectors", "degrain", "adata", "recalc", "compensate"):
        if stage in got:
            assert got[stage] == GOLD_MAIN[case][stage], "%s: plugin %s differs from the reference plugin" % (case, stage)


@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3", "noise_badblocks_blk8"])
def test_flowfps_clip_host_matches_goldens(gpu_ctx, case):
    """mvb200_flowfps_clip_host (Super -> 2 x Analyse -> FlowFPS for a window of output frames in one call, host buffers):
    the whole clip at once, and a window in the middle with exactly the halo the plugin computes"""
    L = mb.lib()
    ctx = gpu_ctx
    frames = pipeline.make_frames(pipeline.CASES[case][0])
    N = len(frames)
    sp, aps, P = pipeline_flow.flow_params_for(case)
    cl = mb.clip_layout(sp)
    host = np.zeros((N, cl.frame_bytes), dtype=np.uint8)
    for n, fr in enumerate(frames):
        for p, a in enumerate(fr):
            rows = host[n, cl.offset[p]: cl.offset[p] + cl.pitch[p] * cl.plane_h[p]].reshape(cl.plane_h[p], cl.pitch[p])
            rows[:, : a.shape[1] * a.itemsize] = a.view(np.uint8).reshape(a.shape[0], -1)
    for i, p in enumerate(P["fps"]):
        M = p.num_frames_out
        out = np.zeros((M, cl.frame_bytes), dtype=np.uint8)
        mb._check(L.mvb200_flowfps_clip_host(ctx.handle, C.byref(sp), C.byref(aps[0]), C.byref(aps[1]), C.byref(p), host.ctypes.data, 0, N, 0, M, out.ctypes.data, None))
        for n in range(M):
            assert [pipeline.sha(q) for q in _planes(out[n], cl)] == GOLD[case]["fps"][i][n], "variant %d frame %d" % (i, n)
        # a window of three output frames from the middle, with the segment the plugin would pass
        n0, n1, off = M // 2, min(M // 2 + 3, M), p.inter.vd.nDeltaFrame
        lmin, lmax = mb.flowfps_position(p, n0)[0], mb.flowfps_position(p, n1 - 1)[0]
        s0, s1 = max(0, lmin - off), min(N, lmax + 2 * off + 1)
        s0 = min(s0, N - 1)
        win = np.zeros((n1 - n0, cl.frame_bytes), dtype=np.uint8)
        mb._check(L.mvb200_flowfps_clip_host(ctx.handle, C.byref(sp), C.byref(aps[0]), C.byref(aps[1]), C.byref(p), host[s0:].ctypes.data, s0, s1 - s0, n0, n1 - n0, win.ctypes.data, None))
        for n in range(n0, n1):
            assert [pipeline.sha(q) for q in _planes(win[n - n0], cl)] == GOLD[case]["fps"][i][n], "variant %d window frame %d" % (i, n)


@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3"])
def test_flowfps_script_through_plugin_with_lookahead(built, monkeypatch, case):
    """MVB200_LOOKAHEAD=4: mv.FlowFPS on the plugin's own Super / Analyse nodes takes the fused path"""
    monkeypatch.setenv("MVB200_LOOKAHEAD", "4")
    got = pipeline_flow.run_reference_flow(case, plugin_path=mb.PLUGIN_PATH, only=("fps",))
    assert got["fps"] == GOLD[case]["fps"], "%s: plugin mv.FlowFPS (fused) differs from the reference plugin" % case


@pytest.mark.parametrize("case", ["c1_640x360_p8_blk8_pel1", "blk16_ov8_pel2_p16_r3", "noise_badblocks_blk8"])
def test_compensate_clip_host_matches_goldens(gpu_ctx, case):
    """mvb200_compensate_clip_host (Super -> Analyse -> Compensate for a window of frames in one call, host buffers):
    the whole clip at once, and an inner window with a one-frame halo"""
    L = mb.lib()
    ctx = gpu_ctx
    frames = pipeline.make_frames(pipeline.CASES[case][0])
    N = len(frames)
    sp, aps, dp = pipeline.params_for(case)
    rp, cp = pipeline.extra_params_for(case)
    cl = mb.clip_layout(sp)
    host = np.zeros((N, cl.frame_bytes), dtype=np.uint8)
    for n, fr in enumerate(frames):
        for p, a in enumerate(fr):
            rows = host[n, cl.offset[p]: cl.offset[p] + cl.pitch[p] * cl.plane_h[p]].reshape(cl.plane_h[p], cl.pitch[p])
            rows[:, : a.shape[1] * a.itemsize] = a.view(np.uint8).reshape(a.shape[0], -1)
    out = np.zeros((N, cl.frame_bytes), dtype=np.uint8)
    mb._check(L.mvb200_compensate_clip_host(ctx.handle, C.byref(sp), C.byref(aps[0]), C.byref(cp), host.ctypes.data, 0, N, 0, N, out.ctypes.data))
    for n in range(N):
        assert [pipeline.sha(q) for q in _planes(out[n], cl)][: cp.nplanes] == GOLD_MAIN[case]["compensate"][n], "frame %d" % n
    win = np.zeros((2, cl.frame_bytes), dtype=np.uint8)
    mb._check(L.mvb200_compensate_clip_host(ctx.handle, C.byref(sp), C.byref(aps[0]), C.byref(cp), host[1:].ctypes.data, 1, 4, 2, 2, win.ctypes.data))
    for n in (2, 3):
        assert [pipeline.sha(q) for q in _planes(win[n - 2], cl)][: cp.nplanes] == GOLD_MAIN[case]["compensate"][n], "window frame %d" % n
