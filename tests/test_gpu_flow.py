"""GPU parity tests (B200) for the consumers of the vector clips: mv.SCDetection / mv.Mask / mv.Finest /
mv.Flow through the C ABI (include/mvtools_b200.h) against the reference's goldens (tests/golden/golden_flow.json) and,
byte for byte, against the oracle; per-frame host-buffer entry points and the device-resident batch entry points."""
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

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_flow.json")))


@pytest.mark.parametrize("case", sorted(c for c in pipeline_flow.FLOW if c not in pipeline.LATE))  # (LATE: test_zz_gpu_pending.py)
def test_cuda_flow_matches_reference_goldens(gpu_ctx, case):
    ctx = gpu_ctx
    n0 = ctx.launch_count()
    got = pipeline_flow.run_engine_flow(case, ctx.scdetection_frame, ctx.mask_frame, ctx.finest_frame, ctx.flow_frame)
    got["inter"] = pipeline_flow.run_engine_inter(case, ctx.flowinter_frame, ctx.blend_frame)
    got["fps"] = pipeline_flow.run_engine_fps(case, ctx.flowinter_frame, ctx.blend_frame)
    assert ctx.launch_count() > n0, "no kernel was launched"
    for stage in ("scd", "mask", "flow", "finest", "inter", "fps"):
        assert got[stage] == GOLD[case][stage], "%s: CUDA %s differs from the reference's golden hashes" % (case, stage)


@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "p10_444_blk8_pel4_bicubic"])
def test_cuda_flow_matches_oracle_bytes(gpu_ctx, case):
    """byte-level comparison against the oracle (locates a mismatch instead of a hash); includes a field shift"""
    import oracle
    ctx = gpu_ctx
    frames, sups, vecs = pipeline_flow.inputs(case)
    N = len(frames)
    sp, aps, P = pipeline_flow.flow_params_for(case)
    for v, fp in P["flow"]:
        for n in range(N):
            r = pipeline.nref_of(aps[v], n, N)
            for fs in (0, 1):
                a = ctx.flow_frame(fp, sp, sups[r] if r is not None else None, vecs[v][n], fs)
                b = oracle.flow_frame(fp, sp, sups[r] if r is not None else None, vecs[v][n], fs)
                assert (a is None) == (b is None)
                if a is not None:
                    for p, (pa, pb) in enumerate(zip(a, b)):
                        bad = np.argwhere(pa != pb)
                        assert bad.size == 0, "flow mode %d frame %d plane %d: first differing sample at %s" % (fp.mode, n, p, bad[0])
    for v, mp in P["mask"]:
        for n in range(N):
            a = ctx.mask_frame(mp, frames[n][0], vecs[v][n])
            b = oracle.mask_frame(mp, frames[n][0], vecs[v][n])
            for p, (pa, pb) in enumerate(zip(a, b)):
                bad = np.argwhere(pa != pb)
                assert bad.size == 0, "mask kind %d frame %d plane %d: first differing sample at %s" % (mp.kind, n, p, bad[0])


@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3", "noise_badblocks_blk8", "p8_420_blk16_ov4_uncovered"])
def test_flow_batch_entry_points_match_goldens(gpu_ctx, case):
    """the device-resident batch entry points (vectors and super clip never leave HBM) give the same bytes"""
    import torch
    L = mb.lib()
    ctx = gpu_ctx
    frames, sups, vecs = pipeline_flow.inputs(case)
    N = len(frames)
    sp, aps, P = pipeline_flow.flow_params_for(case)
    cl, sl = mb.clip_layout(sp), mb.super_layout(sp)
    dsrc = _device_clip(frames, cl)
    dsup = _device_clip(sups, sl)
    vs = (len(vecs[0][0]) + 255) // 256 * 256
    dvec = []
    for v in range(2):
        host = np.zeros((N, vs), dtype=np.uint8)
        for n in range(N):
            host[n, : len(vecs[v][n])] = vecs[v][n]
        dvec.append(torch.from_numpy(host).cuda())
    for i, (v, scp) in enumerate(P["scd"]):
        flags = (C.c_int32 * N)()
        mb._check(L.mvb200_scdetection_batch(ctx.handle, C.byref(scp), dvec[v].data_ptr(), vs, N, flags))
        assert list(flags) == GOLD[case]["scd"][i]
    for i, (v, fp) in enumerate(P["flow"]):
        ddst = torch.full((N, cl.frame_bytes), 0xA5, dtype=torch.uint8, device="cuda")
        mb._check(L.mvb200_flow_batch(ctx.handle, C.byref(fp), C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), 0, N, dvec[v].data_ptr(), vs, ddst.data_ptr()))
        ctx.synchronize()
        out = ddst.cpu().numpy()
        for n in range(N):
            assert [pipeline.sha(p) for p in _planes(out[n], cl)] == GOLD[case]["flow"][i][n], "flow variant %d frame %d" % (i, n)
    for i, (v, mp) in enumerate(P["mask"]):
        ml = mb.mask_layout(mp)
        ddst = torch.full((N, ml.frame_bytes), 0xA5, dtype=torch.uint8, device="cuda")
        mb._check(L.mvb200_mask_batch(ctx.handle, C.byref(mp), dsrc.data_ptr() + cl.offset[0], cl.pitch[0], cl.frame_bytes, dvec[v].data_ptr(), vs, N, ddst.data_ptr()))
        ctx.synchronize()
        out = ddst.cpu().numpy()
        for n in range(N):
            assert [pipeline.sha(p) for p in _planes(out[n], ml)] == GOLD[case]["mask"][i][n], "mask variant %d frame %d" % (i, n)
    for i, p in enumerate(P["inter"]):
        ddst = torch.full((N, cl.frame_bytes), 0xA5, dtype=torch.uint8, device="cuda")
        mb._check(L.mvb200_flowinter_batch(ctx.handle, C.byref(p), C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), 0, N, dvec[0].data_ptr(), dvec[1].data_ptr(), vs, ddst.data_ptr()))
        ctx.synchronize()
        out = ddst.cpu().numpy()
        for n in range(N):
            assert [pipeline.sha(q) for q in _planes(out[n], cl)] == GOLD[case]["inter"][i][n], "inter variant %d frame %d" % (i, n)
    if pipeline_flow.FLOW[case].get("finest"):
        fl = mb.finest_layout(sp)
        ddst = torch.full((N, fl.frame_bytes), 0xA5, dtype=torch.uint8, device="cuda")
        mb._check(L.mvb200_finest_batch(ctx.handle, C.byref(sp), dsup.data_ptr(), N, ddst.data_ptr()))
        ctx.synchronize()
        out = ddst.cpu().numpy()
        for n in range(N):
            assert [pipeline.sha(p) for p in _planes(out[n], fl)] == GOLD[case]["finest"][n], "finest frame %d" % n


def test_flow_full_size_1080p_properties(gpu_ctx):
    """1920x1080 YUV420P8, pel 2, blk 16 / overlap 8 (the bench geometry), vectors from the CUDA search: size-independent
    properties of mv.Flow and mv.Mask.  time = 0 fetches the reference frame itself; a zero vector field (a static clip)
    makes Flow the identity at any time; Finest restricted to the integer positions is the padded level-0 plane."""
    ctx = gpu_ctx
    W, H, N = 1920, 1080, 3
    import clipgen
    frames = clipgen.make_clip(W, H, N, bits=8, scene_cut=False)
    sp = mb.super_params(W, H, 8, 1, 1, pel=2)
    sups = [ctx.super_frame(sp, f) for f in frames]
    ap = mb.analyse_params(sp, N, isb=1, blksize=16, overlap=8)
    vec = ctx.analyse_frame(ap, sp, sups[0], sups[1])
    fp0 = mb.flow_params(sp, ap.ad, N, time=0.0, thscd1=16000, thscd2=255)
    out = ctx.flow_frame(fp0, sp, sups[1], vec)
    assert out is not None
    for p in range(3):
        assert np.array_equal(out[p], frames[1][p]), "time=0 must return the reference frame's own samples"
    # static clip: searching a frame against itself gives zero vectors -> Flow at any time is the identity
    vec0 = ctx.analyse_frame(ap, sp, sups[0], sups[0])
    blob = np.frombuffer(vec0.tobytes(), dtype=np.int32)
    for kw in (dict(time=37.0), dict(time=100.0, mode=1)):
        fp = mb.flow_params(sp, ap.ad, N, thscd1=16000, thscd2=255, **kw)
        out = ctx.flow_frame(fp, sp, sups[0], vec0)
        for p in range(3):
            assert np.array_equal(out[p], frames[0][p])
    # masks of a zero vector field: kind 0 is all zero, kinds 3 / 4 are 128 everywhere
    for kind, val in ((0, 0), (3, 128), (4, 128)):
        mp = mb.mask_params(ap.ad, kind=kind, thscd1=16000, thscd2=255)
        m = ctx.mask_frame(mp, frames[0][0], vec0)
        assert all(int(pl.min()) == val and int(pl.max()) == val for pl in m), kind
    fin = ctx.finest_frame(sp, sups[0])
    g = mb.super_geometry(sp)
    for p in range(3):
        lv = g.lv[p][0]
        assert np.array_equal(fin[p][::2, ::2], sups[0][p][: lv.padded_h, : lv.padded_w])
    assert blob[1] == 1
