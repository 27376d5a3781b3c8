"""Round-2 measurement tool (NOT the bench contract's bench.py): device-resident throughput of the vector-clip consumers
-- mv.Flow / mv.FlowInter / mv.FlowBlur / mv.BlockFPS-core / mv.Mask / mv.Finest -- on a 1080p clip, timed with CUDA events on the
engine's stream, with the algorithmic bytes per frame of DESIGN.md section 4.5 against the measured HBM peak.

    python tools/bench_flow.py [--frames 64] [--bits 8] [--reps 5]

Prints one JSON line per filter.  Inputs are larger than L2 (64 x 17.9 MB super frames)."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vapoursynth-mvtools_b200"))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--bits", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    import clipgen
    import mvtools_b200 as mb
    L = mb.lib()
    W, H, N, bits = 1920, 1080, a.frames, a.bits
    base = clipgen.make_clip(W, H, min(N, 16), bits=bits, scene_cut=False)
    frames = [base[i % len(base)] for i in range(N)]
    sp = mb.super_params(W, H, bits, 1, 1, pel=2)
    cl, sl = mb.clip_layout(sp), mb.super_layout(sp)
    ctx = mb.Context(0)
    host = np.zeros((N, cl.frame_bytes), dtype=np.uint8)
    for n, fr in enumerate(frames):
        for p, pl in enumerate(fr):
            rows = host[n, cl.offset[p]: cl.offset[p] + cl.pitch[p] * cl.plane_h[p]].reshape(cl.plane_h[p], cl.pitch[p])
            rows[:, : pl.shape[1] * pl.itemsize] = pl.view(np.uint8).reshape(pl.shape[0], -1)
    dsrc = torch.from_numpy(host).cuda()
    dsup = torch.empty((N, sl.frame_bytes), dtype=torch.uint8, device="cuda")
    aps = [mb.analyse_params(sp, N, isb=isb, delta=1, blksize=16, overlap=8) for isb in (1, 0)]
    vs = (aps[0].vectors_size + 255) // 256 * 256
    dvec = [torch.zeros((N, vs), dtype=torch.uint8, device="cuda") for _ in aps]
    apa = (C.POINTER(mb.AnalyseParams) * 2)(*[C.pointer(x) for x in aps])
    outs = (C.c_void_p * 2)(*[t.data_ptr() for t in dvec])
    mb._check(L.mvb200_super_batch(ctx.handle, C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), N))
    mb._check(L.mvb200_analyse_multi_batch(ctx.handle, apa, 2, C.byref(sp), dsup.data_ptr(), 0, N, outs, vs))
    ctx.synchronize()
    th = dict(thscd1=16000, thscd2=255)  # keep every frame on the motion-compensated path
    fp = mb.flow_params(sp, aps[0].ad, N, time=50.0, **th)
    fps_ = mb.flow_params(sp, aps[0].ad, N, time=50.0, mode=1, **th)
    ip = mb.flowinter_params(sp, aps[0].ad, aps[1].ad, N, **th)
    bp = mb.flowblur_params(sp, aps[0].ad, aps[1].ad, N, **th)
    mp = mb.mask_params(aps[0].ad, clip_bits=8, kind=2, **th) if bits == 8 else None
    ml = mb.mask_layout(mp) if mp else None
    fl = mb.finest_layout(sp)
    ddst = torch.empty((N, max(cl.frame_bytes, ml.frame_bytes if ml else 0)), dtype=torch.uint8, device="cuda")
    dfin = torch.empty((min(N, 16), fl.frame_bytes), dtype=torch.uint8, device="cuda")
    bps = (bits + 7) // 8
    P = 1.5 * W * H * bps                      # one 4:2:0 frame
    S0 = 4 * 1.5 * (W + 32) * (H + 32) * bps   # level 0 of a pel-2 super frame
    jobs = [
        ("Flow(fetch)", 2 * P, N, lambda: L.mvb200_flow_batch(ctx.handle, C.byref(fp), C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), 0, N, dvec[0].data_ptr(), vs, ddst.data_ptr())),
        ("Flow(shift)", 2 * P + 8 * W * H * 1.5, N, lambda: L.mvb200_flow_batch(ctx.handle, C.byref(fps_), C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), 0, N, dvec[0].data_ptr(), vs, ddst.data_ptr())),
        ("FlowInter", 3 * P, N, lambda: L.mvb200_flowinter_batch(ctx.handle, C.byref(ip), C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), 0, N, dvec[0].data_ptr(), dvec[1].data_ptr(), vs, ddst.data_ptr())),
        ("FlowBlur", 2 * P, N, lambda: L.mvb200_flowblur_batch(ctx.handle, C.byref(bp), C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), 0, N, dvec[0].data_ptr(), dvec[1].data_ptr(), vs, ddst.data_ptr())),
        ("Finest", 2 * S0, min(N, 16), lambda: L.mvb200_finest_batch(ctx.handle, C.byref(sp), dsup.data_ptr(), min(N, 16), dfin.data_ptr())),
    ]
    if mp:
        jobs.append(("Mask(kind=2)", 1.5 * W * H, N, lambda: L.mvb200_mask_batch(ctx.handle, C.byref(mp), dsrc.data_ptr() + cl.offset[0], cl.pitch[0], cl.frame_bytes, dvec[0].data_ptr(), vs, N, ddst.data_ptr())))
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    stream = torch.cuda.ExternalStream(ctx.stream())
    for name, bytes_per_frame, nframes, fn in jobs:
        mb._check(fn())
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            for _ in range(a.reps):
                mb._check(fn())
            e1.record()
        ctx.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        gbs = bytes_per_frame * nframes / (ms * 1e-3) / 1e9
        print(json.dumps({"filter": name, "frames": nframes, "ms": ms, "frames_per_s": nframes / (ms * 1e-3), "algorithmic_GB_s": gbs, "measured_peaks": peak}), flush=True)


if __name__ == "__main__":
    main()
