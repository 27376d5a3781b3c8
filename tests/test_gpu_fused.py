"""GPU parity tests of the fused clip calls of the per-pixel consumers -- mvb200_flow_clip_host, mvb200_flowinter_clip_host,
mvb200_flowblur_clip_host, mvb200_blockfps_clip_host (Super -> Analyse -> consumer for a window of output frames in one
call, host buffers) -- against the reference's goldens, whole clip and inner windows with the plugin's halo, and of the
plugin's fused paths built on them (MVB200_LOOKAHEAD).  All 16 passed on a B200 (profiles/r1_pytest_gpu_fused_clip_calls.log)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import mvtools_b200 as mb
import pipeline
import pipeline_flow
from test_gpu_parity import _planes

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_flow.json")))
CASES = ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3", "noise_badblocks_blk8"]


def _host_clip(case):
    frames = pipeline.make_frames(pipeline.CASES[case][0])
    sp, aps, P = pipeline_flow.flow_params_for(case)
    cl = mb.clip_layout(sp)
    host = np.zeros((len(frames), cl.frame_bytes), dtype=np.uint8)
    for n, fr in enumerate(frames):
        for p, a in enumerate(fr):
            rows = host[n, cl.offset[p]: cl.offset[p] + cl.pitch[p] * cl.plane_h[p]].reshape(cl.plane_h[p], cl.pitch[p])
            rows[:, : a.shape[1] * a.itemsize] = a.view(np.uint8).reshape(a.shape[0], -1)
    return host, sp, aps, P, cl


def _check_frames(out, cl, gold_rows, first, what):
    for k in range(out.shape[0]):
        assert [pipeline.sha(q) for q in _planes(out[k], cl)] == gold_rows[first + k], "%s frame %d" % (what, first + k)


@pytest.mark.parametrize("case", CASES)
def test_flow_clip_host_matches_goldens(gpu_ctx, case):
    """mvb200_flow_clip_host (Super -> Analyse -> Flow in one call, host buffers): the whole clip, and an inner window with
    the one-delta halo the plugin passes; the usable flags say which frames are copies of the source frame"""
    L = mb.lib()
    host, sp, aps, P, cl = _host_clip(case)
    N = host.shape[0]
    for i, (v, fp) in enumerate(P["flow"]):
        if fp.fields:
            continue  # (field shifts: node-by-node evaluation only)
        out = np.zeros((N, cl.frame_bytes), dtype=np.uint8)
        usable = (C.c_int32 * N)()
        mb._check(L.mvb200_flow_clip_host(gpu_ctx.handle, C.byref(sp), C.byref(aps[v]), C.byref(fp), host.ctypes.data, 0, N, 0, N, out.ctypes.data, usable))
        _check_frames(out, cl, GOLD[case]["flow"][i], 0, "variant %d" % i)
        for n in range(N):
            if pipeline.nref_of(aps[v], n, N) is None:
                assert usable[n] == 0, "variant %d frame %d has no reference frame" % (i, n)
            if not usable[n]:
                assert all(np.array_equal(a, b) for a, b in zip(_planes(out[n], cl), _planes(host[n], cl)))  # (pitch padding is not compared)
        d = fp.vd.nDeltaFrame
        n0, n1 = 2, min(4, N)
        s0, s1 = max(0, n0 - d), min(N, n1 + d)
        win = np.zeros((n1 - n0, cl.frame_bytes), dtype=np.uint8)
        mb._check(L.mvb200_flow_clip_host(gpu_ctx.handle, C.byref(sp), C.byref(aps[v]), C.byref(fp), host[s0:].ctypes.data, s0, s1 - s0, n0, n1 - n0, win.ctypes.data, None))
        _check_frames(win, cl, GOLD[case]["flow"][i], n0, "variant %d window" % i)


@pytest.mark.parametrize("case", CASES)
def test_flowinter_clip_host_matches_goldens(gpu_ctx, case):
    """mvb200_flowinter_clip_host (Super -> 2 x Analyse -> FlowInter): whole clip; inner window with delta before and
    2 x delta after it"""
    L = mb.lib()
    host, sp, aps, P, cl = _host_clip(case)
    N = host.shape[0]
    for i, p in enumerate(P["inter"]):
        out = np.zeros((N, cl.frame_bytes), dtype=np.uint8)
        usable = (C.c_int32 * N)()
        mb._check(L.mvb200_flowinter_clip_host(gpu_ctx.handle, C.byref(sp), C.byref(aps[0]), C.byref(aps[1]), C.byref(p), host.ctypes.data, 0, N, 0, N, out.ctypes.data, usable))
        _check_frames(out, cl, GOLD[case]["inter"][i], 0, "variant %d" % i)
        assert all(usable[n] == 0 for n in range(max(0, N - p.vd.nDeltaFrame), N)), "frames without a right neighbour are not interpolated"
        d = p.vd.nDeltaFrame
        n0, n1 = 1, min(3, N)
        s0, s1 = max(0, n0 - d), min(N, n1 + 2 * d)
        win = np.zeros((n1 - n0, cl.frame_bytes), dtype=np.uint8)
        mb._check(L.mvb200_flowinter_clip_host(gpu_ctx.handle, C.byref(sp), C.byref(aps[0]), C.byref(aps[1]), C.byref(p), host[s0:].ctypes.data, s0, s1 - s0, n0, n1 - n0, win.ctypes.data, None))
        _check_frames(win, cl, GOLD[case]["inter"][i], n0, "variant %d window" % i)


@pytest.mark.parametrize("case", CASES)
def test_flowblur_clip_host_matches_goldens(gpu_ctx, case):
    """mvb200_flowblur_clip_host (Super -> 2 x Analyse -> FlowBlur): whole clip; inner window with a delta halo"""
    L = mb.lib()
    host, sp, aps, P, cl = _host_clip(case)
    N = host.shape[0]
    for i, p in enumerate(P["blur"]):
        out = np.zeros((N, cl.frame_bytes), dtype=np.uint8)
        mb._check(L.mvb200_flowblur_clip_host(gpu_ctx.handle, C.byref(sp), C.byref(aps[0]), C.byref(aps[1]), C.byref(p), host.ctypes.data, 0, N, 0, N, out.ctypes.data))
        _check_frames(out, cl, GOLD[case]["blur"][i], 0, "variant %d" % i)
        d = p.vd.nDeltaFrame
        n0, n1 = 2, min(4, N)
        s0, s1 = max(0, n0 - d), min(N, n1 + d)
        win = np.zeros((n1 - n0, cl.frame_bytes), dtype=np.uint8)
        mb._check(L.mvb200_flowblur_clip_host(gpu_ctx.handle, C.byref(sp), C.byref(aps[0]), C.byref(aps[1]), C.byref(p), host[s0:].ctypes.data, s0, s1 - s0, n0, n1 - n0, win.ctypes.data))
        _check_frames(win, cl, GOLD[case]["blur"][i], n0, "variant %d window" % i)


@pytest.mark.parametrize("case", CASES)
def test_blockfps_clip_host_matches_goldens(gpu_ctx, case):
    """mvb200_blockfps_clip_host (Super -> 2 x Analyse -> BlockFPS for a window of OUTPUT frames): the whole clip, and a
    window from the middle with the segment the plugin passes (its frames' left and right source frames)"""
    L = mb.lib()
    host, sp, aps, P, cl = _host_clip(case)
    N = host.shape[0]
    for i, p in enumerate(P["bfps"]):
        M = p.num_frames_out
        out = np.zeros((M, cl.frame_bytes), dtype=np.uint8)
        usable = (C.c_int32 * M)()
        mb._check(L.mvb200_blockfps_clip_host(gpu_ctx.handle, C.byref(sp), C.byref(aps[0]), C.byref(aps[1]), C.byref(p), host.ctypes.data, 0, N, 0, M, out.ctypes.data, usable))
        _check_frames(out, cl, GOLD[case]["bfps"][i], 0, "variant %d" % i)
        for n in range(M):
            nleft, time256 = mb.blockfps_position(p, n)
            if time256 in (0, 256) or nleft + p.vd.nDeltaFrame >= N:
                assert usable[n] == 0, "variant %d frame %d is an input frame / has no right frame" % (i, n)
        n0, n1, off = M // 2, min(M // 2 + 3, M), p.vd.nDeltaFrame
        lmin, lmax = mb.blockfps_position(p, n0)[0], mb.blockfps_position(p, n1 - 1)[0]
        s0, s1 = min(lmin, N - 1), min(N, lmax + off + 1)
        win = np.zeros((n1 - n0, cl.frame_bytes), dtype=np.uint8)
        mb._check(L.mvb200_blockfps_clip_host(gpu_ctx.handle, C.byref(sp), C.byref(aps[0]), C.byref(aps[1]), C.byref(p), host[s0:].ctypes.data, s0, s1 - s0, n0, n1 - n0, win.ctypes.data, None))
        _check_frames(win, cl, GOLD[case]["bfps"][i], n0, "variant %d window" % i)


@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3"])
def test_blockfps_script_through_plugin_with_lookahead(built, monkeypatch, case):
    """MVB200_LOOKAHEAD=4: mv.BlockFPS on the plugin's own Super / Analyse nodes takes the fused path"""
    monkeypatch.setenv("MVB200_LOOKAHEAD", "4")
    got = pipeline_flow.run_reference_flow(case, plugin_path=mb.PLUGIN_PATH, only=("bfps",))
    assert got["bfps"] == GOLD[case]["bfps"], "%s: plugin mv.BlockFPS (fused) differs from the reference plugin" % case


@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3"])
def test_flow_scripts_through_plugin_with_lookahead(built, monkeypatch, case):
    """MVB200_LOOKAHEAD=4: mv.Flow / mv.FlowInter / mv.FlowBlur on the plugin's own Super / Analyse nodes take the fused path"""
    monkeypatch.setenv("MVB200_LOOKAHEAD", "4")
    got = pipeline_flow.run_reference_flow(case, plugin_path=mb.PLUGIN_PATH, only=("flow", "inter", "blur"))
    for stage in ("flow", "inter", "blur"):
        assert got[stage] == GOLD[case][stage], "%s: plugin mv.%s (fused) differs from the reference plugin" % (case, stage)
