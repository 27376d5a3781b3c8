"""GPU tests that have NOT run on a GPU yet: (1) the block shapes / penalty options of pipeline.LATE (64x32, 64x64, 128x64,
128x128, 16x2, 16x8, 32x16 blocks, search 7, explicit lambda / pnew / plevel / pzero / pglobal / lsad / levels) through
the per-frame entry points -- the oracle is pinned against the reference on them; (2) the fused clip calls of the per-pixel consumers (mvb200_flow_clip_host,
mvb200_flowinter_clip_host, mvb200_flowblur_clip_host, mvb200_blockfps_clip_host) and the plugin's fused paths built on them were written after
round 1's GPU minutes were spent.  They are compositions of pieces that HAVE passed on a B200 (mvb200_super_batch,
mvb200_analyse_multi_batch, mvb200_flow_batch / _flowinter_batch / _flowblur_batch -- the same composition as
mvb200_flowfps_clip_host, test_gpu_flow2.py) and the plugin glue around them is tested on the CPU against a stand-in
engine (test_plugin_glue.py); what remains unproven is these entry points' own buffer arithmetic on the device.  The file
sorts last so that `pytest -x` reaches it after every test that has already passed on a B200."""
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
GOLD_MAIN = json.load(open(os.path.join(HERE, "golden", "golden.json")))
CASES = ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3", "noise_badblocks_blk8"]


@pytest.mark.parametrize("case", sorted(pipeline.LATE))
def test_cuda_matches_reference_goldens_late_cases(gpu_ctx, case):
    """per-frame host-buffer entry points on the LATE cases: bit-exact against the reference's goldens"""
    ctx = gpu_ctx
    got = pipeline.run_engine(case, ctx.super_frame, ctx.analyse_frame, ctx.degrain_frame, ctx.recalculate_frame, ctx.compensate_frame)
    for stage in ("super", "vectors", "degrain", "recalc", "compensate"):
        if stage in got:
            assert got[stage] == GOLD_MAIN[case][stage], "%s: CUDA %s differs from the reference" % (case, stage)


@pytest.mark.parametrize("case", sorted(c for c in pipeline_flow.FLOW if c in pipeline.LATE))
def test_cuda_flow_group_matches_reference_goldens_late_cases(gpu_ctx, case):
    """the vector-clip consumers on the LATE cases (64x64, 64x32, 16x8, 128x64 blocks): per-frame entry points against the
    reference's goldens (the kernels' per-thread code already reproduces them in the CPU emulation, test_flow_emul.py)"""
    ctx = gpu_ctx
    got = pipeline_flow.run_engine_flow(case, ctx.scdetection_frame, ctx.mask_frame, ctx.finest_frame, ctx.flow_frame)
    got["inter"] = pipeline_flow.run_engine_inter(case, ctx.flowinter_frame, ctx.blend_frame)
    got["fps"] = pipeline_flow.run_engine_fps(case, ctx.flowinter_frame, ctx.blend_frame)
    got["blur"] = pipeline_flow.run_engine_blur(case, ctx.flowblur_frame)
    got["bfps"] = pipeline_flow.run_engine_bfps(case, ctx.blockfps_frame, ctx.blend_frame)
    for stage in ("scd", "mask", "flow", "finest", "inter", "fps", "blur", "bfps"):
        assert got[stage] == GOLD[case][stage], "%s: CUDA %s differs from the reference's golden hashes" % (case, stage)


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
