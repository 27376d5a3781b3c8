"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against
 (1) the CPU oracle on the same seeded inputs, (2) the committed golden hashes of the
 reference's own outputs, and (3) at BASELINE.json's full 1080p size, through
 size-independent properties plus a sampled oracle comparison."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import mvtools_b200 as mb
import pipeline

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


@pytest.mark.parametrize("case", sorted(c for c in pipeline.CASES if c not in pipeline.LATE))  # (LATE: test_zz_gpu_pending.py)
def test_cuda_matches_reference_goldens(gpu_ctx, case):
    """per-frame host-buffer entry points (what the VapourSynth plugin calls): bit-exact"""
    ctx = gpu_ctx
    got = pipeline.run_engine(case, ctx.super_frame, ctx.analyse_frame, ctx.degrain_frame, ctx.recalculate_frame, ctx.compensate_frame)
    for stage in ("super", "vectors", "degrain", "recalc", "compensate"):
        if stage in got:
            assert got[stage] == GOLD[case][stage], "%s: CUDA %s differs from the reference" % (case, stage)


@pytest.mark.parametrize("case", ["c1_640x360_p8_blk8_pel1", "blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3", "blk16_ov8_pel2_p8_r3",
                                  "noise_badblocks_blk8", "p10_444_blk8_pel4_bicubic", "blk32_p16_umh"])
def test_generic_kernel_matches_reference_goldens(gpu_ctx, case, monkeypatch):
    """the same cases forced through the generic (fallback) search kernel instead of the specialised one"""
    monkeypatch.setenv("MVB200_FORCE_GENERIC", "1")
    ctx = gpu_ctx
    got = pipeline.run_engine(case, ctx.super_frame, ctx.analyse_frame, ctx.degrain_frame)
    for stage in ("super", "vectors", "degrain"):
        assert got[stage] == GOLD[case][stage], "%s: CUDA (generic kernel) %s differs from the reference" % (case, stage)


@pytest.mark.parametrize("case", ["blk8_ov4_pel2_p8", "blk16_ov8_pel2_p16_r3", "noise_badblocks_blk8"])
def test_cuda_matches_oracle_bytes(gpu_ctx, case):
    """byte-level comparison against the oracle (locates a mismatch instead of a hash)"""
    import oracle
    ctx = gpu_ctx
    frames = pipeline.make_frames(pipeline.CASES[case][0])
    N = len(frames)
    sp, aps, dp = pipeline.params_for(case)
    gs = [ctx.super_frame(sp, f) for f in frames]
    os_ = [oracle.super_frame(sp, f) for f in frames]
    for a, b in zip(gs, os_):
        for pa, pb in zip(a, b):
            assert np.array_equal(pa, pb)
    for n in range(N):
        gv, ov = [], []
        for ap in aps:
            r = pipeline.nref_of(ap, n, N)
            gv.append(ctx.analyse_frame(ap, sp, gs[n], gs[r] if r is not None else None))
            ov.append(oracle.analyse_frame(ap, sp, os_[n], os_[r] if r is not None else None))
            bad = np.flatnonzero(gv[-1] != ov[-1])
            assert bad.size == 0, "frame %d: first differing blob byte %d" % (n, bad[0])
        refs = [gs[pipeline.nref_of(ap, n, N)] if pipeline.nref_of(ap, n, N) is not None else None for ap in aps]
        gd = ctx.degrain_frame(dp, sp, frames[n], refs, gv)
        od = oracle.degrain_frame(dp, sp, frames[n], refs, ov)
        for pa, pb in zip(gd, od):
            assert np.array_equal(pa, pb)


def _device_clip(frames, lay):
    import torch
    N = len(frames)
    host = np.zeros((N, lay.frame_bytes), dtype=np.uint8)
    for n, fr in enumerate(frames):
        for p, a in enumerate(fr):
            rows = host[n, lay.offset[p]: lay.offset[p] + lay.pitch[p] * lay.plane_h[p]].reshape(lay.plane_h[p], lay.pitch[p])
            rows[:, : a.shape[1] * a.itemsize] = a.view(np.uint8).reshape(a.shape[0], -1)
    return torch.from_numpy(host).cuda()


def _planes(buf, lay):
    out = []
    for p in range(lay.nplanes):
        a = buf[lay.offset[p]: lay.offset[p] + lay.pitch[p] * lay.plane_h[p]].reshape(lay.plane_h[p], lay.pitch[p])[:, : lay.plane_w[p] * lay.bytes_per_sample]
        a = np.ascontiguousarray(a)
        out.append(a.view(np.uint16) if lay.bytes_per_sample == 2 else a)
    return out


def _run_batch(ctx, sp, aps, dp, frames):
    import torch
    L = mb.lib()
    N = len(frames)
    cl, sl = mb.clip_layout(sp), mb.super_layout(sp)
    dsrc = _device_clip(frames, cl)
    dsup = torch.empty((N, sl.frame_bytes), dtype=torch.uint8, device="cuda")
    vs = (aps[0].vectors_size + 255) // 256 * 256
    dvec = [torch.zeros((N, vs), dtype=torch.uint8, device="cuda") for _ in aps]
    ddst = torch.empty((N, cl.frame_bytes), dtype=torch.uint8, device="cuda")
    apa = (C.POINTER(mb.AnalyseParams) * len(aps))(*[C.pointer(a) for a in aps])
    outs = (C.c_void_p * len(aps))(*[t.data_ptr() for t in dvec])
    mb._check(L.mvb200_super_batch(ctx.handle, C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), N))
    mb._check(L.mvb200_analyse_multi_batch(ctx.handle, apa, len(aps), C.byref(sp), dsup.data_ptr(), 0, N, outs, vs))
    mb._check(L.mvb200_degrain_batch(ctx.handle, C.byref(dp), C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), N, 0, N, outs, vs, ddst.data_ptr()))
    ctx.synchronize()
    return dsup.cpu().numpy(), [d.cpu().numpy() for d in dvec], ddst.cpu().numpy(), cl, sl


@pytest.mark.parametrize("case", ["c1_640x360_p8_blk8_pel1", "blk16_ov8_pel2_p8_r3", "blk16_ov8_pel2_p16_r3", "p8_422_blk8x4_sharp0_rf1"])
def test_batch_path_matches_goldens(gpu_ctx, case):
    """the device-resident batch entry points (what bench.py times) give the same bytes"""
    frames = pipeline.make_frames(pipeline.CASES[case][0])
    sp, aps, dp = pipeline.params_for(case)
    sup, vecs, dst, cl, sl = _run_batch(gpu_ctx, sp, aps, dp, frames)
    for n in range(len(frames)):
        assert [pipeline.sha(p) for p in _planes(sup[n], sl)] == GOLD[case]["super"][n]
        assert [pipeline.sha(pipeline.mask_blob(v[n][: ap.vectors_size], ap)) for v, ap in zip(vecs, aps)] == GOLD[case]["vectors"][n]
        assert [pipeline.sha(p) for p in _planes(dst[n], cl)] == GOLD[case]["degrain"][n]


@pytest.mark.parametrize("case", ["blk16_ov8_pel2_p8_r3", "p8_422_blk8x4_sharp0_rf1"])
def test_host_clip_pipeline_matches_goldens(gpu_ctx, case):
    """mvb200_degrain_clip_host_async: host buffers in, host buffers out, three overlapping calls (both
    device slots reused) -- every call's output equals the goldens"""
    import torch
    L = mb.lib()
    frames = pipeline.make_frames(pipeline.CASES[case][0])
    sp, aps, dp = pipeline.params_for(case)
    N = len(frames)
    cl = mb.clip_layout(sp)
    host = torch.zeros((N, cl.frame_bytes), dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    for n, fr in enumerate(frames):
        for p, a in enumerate(fr):
            rows = hv[n, cl.offset[p]: cl.offset[p] + cl.pitch[p] * cl.plane_h[p]].reshape(cl.plane_h[p], cl.pitch[p])
            rows[:, : a.shape[1] * a.itemsize] = a.view(np.uint8).reshape(a.shape[0], -1)
    outs = [torch.zeros((N, cl.frame_bytes), dtype=torch.uint8).pin_memory() for _ in range(3)]
    apa = (C.POINTER(mb.AnalyseParams) * len(aps))(*[C.pointer(a) for a in aps])
    ranges = [(0, N), (0, N), (1, N - 2)]  # the last call denoises an inner range only (halo frames as references)
    for o, (first, count) in zip(outs, ranges):
        mb._check(L.mvb200_degrain_clip_host_async(gpu_ctx.handle, C.byref(sp), apa, len(aps), C.byref(dp), host.data_ptr(), o.data_ptr(), N, first, count))
    gpu_ctx.synchronize()
    for o, (first, count) in zip(outs, ranges):
        dst = o.numpy()
        for k in range(count):
            assert [pipeline.sha(p) for p in _planes(dst[k], cl)] == GOLD[case]["degrain"][first + k]


@pytest.mark.parametrize("case", sorted(c for c in pipeline.EXTRA if c not in pipeline.LATE))
def test_recalculate_compensate_batch_match_goldens(gpu_ctx, case):
    """mvb200_recalculate_batch / mvb200_compensate_batch on a device-resident clip (vector clip 0 of the case, as in
    the per-frame test) give the reference's bytes for every frame"""
    import torch
    L = mb.lib()
    ctx = gpu_ctx
    if case in pipeline.PELCLIP:
        pytest.skip("pelclip cases go through the per-frame Super entry point")
    frames = pipeline.make_frames(pipeline.CASES[case][0])
    N = len(frames)
    sp, aps, dp = pipeline.params_for(case)
    rp, cp = pipeline.extra_params_for(case)
    cl, sl = mb.clip_layout(sp), mb.super_layout(sp)
    dsrc = _device_clip(frames, cl)
    dsup = torch.empty((N, sl.frame_bytes), dtype=torch.uint8, device="cuda")
    vs = (aps[0].vectors_size + 255) // 256 * 256
    dvec = torch.zeros((N, vs), dtype=torch.uint8, device="cuda")
    rs = (rp.vectors_size + 255) // 256 * 256
    dnew = torch.zeros((N, rs), dtype=torch.uint8, device="cuda")
    dcomp = torch.full((N, cl.frame_bytes), 0xA5, dtype=torch.uint8, device="cuda")
    apa = (C.POINTER(mb.AnalyseParams) * 1)(C.pointer(aps[0]))
    outs = (C.c_void_p * 1)(dvec.data_ptr())
    mb._check(L.mvb200_super_batch(ctx.handle, C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), N))
    mb._check(L.mvb200_analyse_multi_batch(ctx.handle, apa, 1, C.byref(sp), dsup.data_ptr(), 0, N, outs, vs))
    mb._check(L.mvb200_recalculate_batch(ctx.handle, C.byref(rp), C.byref(sp), dsup.data_ptr(), 0, N, dvec.data_ptr(), vs, dnew.data_ptr(), rs))
    mb._check(L.mvb200_compensate_batch(ctx.handle, C.byref(cp), C.byref(sp), dsup.data_ptr(), 0, N, dvec.data_ptr(), vs, dcomp.data_ptr()))
    ctx.synchronize()
    new, comp = dnew.cpu().numpy(), dcomp.cpu().numpy()
    for n in range(N):
        assert pipeline.sha(pipeline.mask_recalc_blob(new[n][: rp.vectors_size], rp)) == GOLD[case]["recalc"][n], "frame %d: Recalculate" % n
        assert [pipeline.sha(p) for p in _planes(comp[n], cl)][: cp.nplanes] == GOLD[case]["compensate"][n], "frame %d: Compensate" % n
    # an inner range addresses the same frames
    dnew2 = torch.zeros((N, rs), dtype=torch.uint8, device="cuda")
    mb._check(L.mvb200_recalculate_batch(ctx.handle, C.byref(rp), C.byref(sp), dsup.data_ptr(), 1, N - 2, dvec[1:].data_ptr(), vs, dnew2.data_ptr(), rs))
    ctx.synchronize()
    assert np.array_equal(dnew2.cpu().numpy()[: N - 2, : rp.vectors_size], new[1: N - 1, : rp.vectors_size])



def test_host_clip_pipeline_pieces_match_batch_path(gpu_ctx):
    """a segment long enough to move in pieces (kClipChunks x 8 frames or more: chunked H2D + Super, chunked
    Degrain + D2H) gives the bytes of the device-resident batch path, for a full and for an inner output range"""
    import torch
    L = mb.lib()
    case = "blk16_ov8_pel2_p8_r3"
    ck = dict(pipeline.CASES[case][0], nframes=38)
    frames = pipeline.make_frames(ck)
    N = len(frames)
    W, H = ck["width"], ck["height"]
    _, skw, akw, radius, dkw = pipeline.CASES[case]
    sp = mb.super_params(W, H, ck["bits"], 1, 1, False, **skw)
    aps = [mb.analyse_params(sp, N, isb=isb, delta=d, **akw) for d in range(1, radius + 1) for isb in (1, 0)]
    dp = mb.degrain_params(sp, [ap.ad for ap in aps], W, H, radius, **dkw)
    _, _, want, cl, _ = _run_batch(gpu_ctx, sp, aps, dp, frames)
    host = torch.zeros((N, cl.frame_bytes), dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    for n, fr in enumerate(frames):
        for p, a in enumerate(fr):
            rows = hv[n, cl.offset[p]: cl.offset[p] + cl.pitch[p] * cl.plane_h[p]].reshape(cl.plane_h[p], cl.pitch[p])
            rows[:, : a.shape[1] * a.itemsize] = a.view(np.uint8).reshape(a.shape[0], -1)
    outs = [torch.zeros((N, cl.frame_bytes), dtype=torch.uint8).pin_memory() for _ in range(3)]
    apa = (C.POINTER(mb.AnalyseParams) * len(aps))(*[C.pointer(a) for a in aps])
    ranges = [(0, N), (3, N - 6), (0, N)]
    for o, (first, count) in zip(outs, ranges):
        mb._check(L.mvb200_degrain_clip_host_async(gpu_ctx.handle, C.byref(sp), apa, len(aps), C.byref(dp), host.data_ptr(), o.data_ptr(), N, first, count))
    gpu_ctx.synchronize()
    for o, (first, count) in zip(outs, ranges):
        dst = o.numpy()
        for k in range(count):
            for a, b in zip(_planes(dst[k], cl), _planes(want[first + k], cl)):
                assert np.array_equal(a, b), "frame %d of range (%d, %d)" % (first + k, first, count)



def test_full_size_1080p_properties(gpu_ctx):
    """BASELINE.json configs[1]-sized run (1920x1080 P8, blk 8 ov 4 pel 2, Degrain1) checked by
    properties that do not need the slow CPU path at full size, plus one oracle frame."""
    import clipgen
    import oracle
    W, H, N = 1920, 1080, 5
    base = clipgen.make_clip(W, H, 3, bits=8, scene_cut=False)
    frames = [base[0], base[1], base[2], base[2], base[1]]  # frame 3 == frame 2 (static pair)
    sp = mb.super_params(W, H, 8, 1, 1, pel=2)
    aps = [mb.analyse_params(sp, N, isb=isb, blksize=8, overlap=4) for isb in (1, 0)]
    dp = mb.degrain_params(sp, [a.ad for a in aps], W, H, 1)
    sup, vecs, dst, cl, sl = _run_batch(gpu_ctx, sp, aps, dp, frames)
    g = mb.super_geometry(sp)
    # (1) level 0 / sub-plane 0 interior of the super frame is the source frame
    for p in range(3):
        lv = g.lv[p][0]
        pl = _planes(sup[1], sl)[p]
        assert np.array_equal(pl[lv.vpad: lv.vpad + lv.height, lv.hpad: lv.hpad + lv.width], frames[1][p])
    # (2) identical frames give zero vectors with zero SAD at the finest level (frame 2 backward -> 3)
    ap = aps[0]
    blob = vecs[0][2][: ap.vectors_size]
    n0 = ap.nBlkX[0] * ap.nBlkY[0]
    lvl0 = np.frombuffer(blob[ap.vectors_size - n0 * 16:].tobytes(), dtype=np.int32).reshape(n0, 4)
    assert not lvl0.any(), "static frame pair must give all-zero vectors and SADs"
    assert np.frombuffer(blob[:8].tobytes(), dtype=np.int32).tolist() == [ap.vectors_size, 1]
    # (3) blob structure: level size headers add up; out-of-range reference -> validity 0
    assert np.frombuffer(vecs[1][0][:8].tobytes(), dtype=np.int32).tolist() == [ap.vectors_size, 0]
    off = 8
    for l in range(ap.ad.nLvCount - 1, -1, -1):
        sz = int(np.frombuffer(blob[off:off + 4].tobytes(), dtype=np.int32)[0])
        assert sz == 4 + 16 * ap.nBlkX[l] * ap.nBlkY[l]
        off += sz
    assert off == ap.vectors_size
    # (4) the uncovered right/bottom strips of the output equal the source
    out = _planes(dst[1], cl)
    assert np.array_equal(out[0][:, dp.nWidth_B[0]:], frames[1][0][:, dp.nWidth_B[0]:])
    assert np.array_equal(out[0][dp.nHeight_B[0]:, :], frames[1][0][dp.nHeight_B[0]:, :])
    # (5) one full frame against the oracle at full size (a few seconds of CPU)
    osup = [oracle.super_frame(sp, f) for f in frames[0:3]]
    for p in range(3):
        assert np.array_equal(_planes(sup[1], sl)[p], osup[1][p])
    ov = [oracle.analyse_frame(aps[0], sp, osup[1], osup[2]), oracle.analyse_frame(aps[1], sp, osup[1], osup[0])]
    for i in range(2):
        assert np.array_equal(vecs[i][1][: ap.vectors_size], ov[i])
    od = oracle.degrain_frame(dp, sp, frames[1], [osup[2], osup[0]], ov)
    for p in range(3):
        assert np.array_equal(out[p], od[p])


def test_full_size_4k_p16_degrain3_properties(gpu_ctx):
    """BASELINE.json configs[4] geometry (3840x2160 P16, blk 16 ov 8, Degrain3) on a short clip: batch path,
    properties, and one frame of every stage against the oracle at full size."""
    import clipgen
    import oracle
    W, H, N = 3840, 2160, 7
    base = clipgen.make_clip(W, H, 4, bits=16, scene_cut=False)
    frames = [base[0], base[1], base[2], base[3], base[3], base[2], base[1]]  # frames 3, 4 identical
    sp = mb.super_params(W, H, 16, 1, 1, pel=2)
    aps = [mb.analyse_params(sp, N, isb=isb, delta=d, blksize=16, overlap=8) for d in (1, 2, 3) for isb in (1, 0)]
    dp = mb.degrain_params(sp, [a.ad for a in aps], W, H, 3)
    sup, vecs, dst, cl, sl = _run_batch(gpu_ctx, sp, aps, dp, frames)
    ap = aps[0]
    n0 = ap.nBlkX[0] * ap.nBlkY[0]
    assert (ap.nBlkX[0], ap.nBlkY[0]) == (479, 269)
    # static pair: frame 3 backward (-> 4) has all-zero vectors and SADs at the finest level
    blob = vecs[0][3][: ap.vectors_size]
    lvl0 = np.frombuffer(blob[ap.vectors_size - n0 * 16:].tobytes(), dtype=np.int32).reshape(n0, 4)
    assert not lvl0.any()
    # validity: delta-3 forward vectors of frames 0..2 have no reference
    for n in range(N):
        want = 1 if n >= 3 else 0
        assert np.frombuffer(vecs[5][n][:8].tobytes(), dtype=np.int32).tolist() == [ap.vectors_size, want]
    # one frame of each stage against the oracle (frame 3: all six references exist)
    n = 3
    osup = {k: oracle.super_frame(sp, frames[k]) for k in range(N)}
    for p in range(3):
        assert np.array_equal(_planes(sup[n], sl)[p], osup[n][p])
    ov = []
    for i, a in enumerate(aps):
        r = pipeline.nref_of(a, n, N)
        ov.append(oracle.analyse_frame(a, sp, osup[n], osup[r]))
        assert np.array_equal(vecs[i][n][: a.vectors_size], ov[i]), "vectors of instance %d" % i
    od = oracle.degrain_frame(dp, sp, frames[n], [osup[pipeline.nref_of(a, n, N)] for a in aps], ov)
    out = _planes(dst[n], cl)
    for p in range(3):
        assert np.array_equal(out[p], od[p])


def test_full_size_4k_satd_recalculate_compensate(gpu_ctx):
    """BASELINE.json configs[3] (3840x2160 P8: Analyse with SATD -- dct=5 in the reference's numbering --, Recalculate,
    Compensate) through the per-frame entry points, against the oracle at full size."""
    import clipgen
    import oracle
    W, H, N = 3840, 2160, 2
    frames = clipgen.make_clip(W, H, N, bits=8, scene_cut=False)
    sp = mb.super_params(W, H, 8, 1, 1, pel=2)
    ap = mb.analyse_params(sp, N, isb=1, delta=1, dct=5)
    rp = mb.recalculate_params(sp, ap.ad, N, blksize=8, overlap=4, thsad=120, dct=5)
    cp = mb.compensate_params(sp, rp.ad, W, H, N)
    gs = [gpu_ctx.super_frame(sp, f) for f in frames]
    os_ = [oracle.super_frame(sp, f) for f in frames]
    for a, b in zip(gs[0], os_[0]):
        assert np.array_equal(a, b)
    gv = gpu_ctx.analyse_frame(ap, sp, gs[0], gs[1])
    ov = oracle.analyse_frame(ap, sp, os_[0], os_[1])
    assert np.array_equal(gv, ov)
    gr = gpu_ctx.recalculate_frame(rp, sp, gs[0], gs[1], gv)
    orr = oracle.recalculate_frame(rp, sp, os_[0], os_[1], ov)
    assert np.array_equal(pipeline.mask_recalc_blob(gr, rp), pipeline.mask_recalc_blob(orr, rp))
    gc = gpu_ctx.compensate_frame(cp, sp, gs[0], gs[1], gr)
    oc = oracle.compensate_frame(cp, sp, os_[0], os_[1], orr)
    for a, b in zip(gc, oc):
        assert np.array_equal(a, b)


def test_smoke_entry(built):
    import __graft_entry__
    __graft_entry__.smoke()
