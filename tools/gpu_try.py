"""Scratch GPU experiment: batch path timing of Super / Analyse xK / DegrainN (not a bench)."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("vapoursynth-mvtools_b200", "oracle", "host/minivs"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
import clipgen, mvtools_b200 as mb

def run(W, H, N, bits, radius, blk, ov, pel, check=True, distinct=8):
    L = mb.lib()
    frames = clipgen.make_clip(W, H, min(N, distinct), bits=bits)
    frames = [frames[i % len(frames)] for i in range(N)]
    sp = mb.super_params(W, H, bits, 1, 1, pel=pel)
    cl, sl = mb.clip_layout(sp), mb.super_layout(sp)
    ctx = mb.Context(0)
    host = np.zeros((N, cl.frame_bytes), dtype=np.uint8)
    for n, fr in enumerate(frames):
        for p, a in enumerate(fr):
            rows = host[n, cl.offset[p]: cl.offset[p] + cl.pitch[p] * cl.plane_h[p]].reshape(cl.plane_h[p], cl.pitch[p])
            rows[:, : a.shape[1] * a.itemsize] = a.view(np.uint8).reshape(a.shape[0], -1)
    dsrc = torch.from_numpy(host).cuda()
    dsup = torch.empty((N, sl.frame_bytes), dtype=torch.uint8, device="cuda")
    aps = [mb.analyse_params(sp, N, isb=isb, delta=d, blksize=blk, overlap=ov) for d in range(1, radius + 1) for isb in (1, 0)]
    vs = (aps[0].vectors_size + 255) // 256 * 256
    dvec = [torch.empty((N, vs), dtype=torch.uint8, device="cuda") for _ in aps]
    dp = mb.degrain_params(sp, [ap.ad for ap in aps], W, H, radius)
    ddst = torch.empty((N, cl.frame_bytes), dtype=torch.uint8, device="cuda")
    apa = (C.POINTER(mb.AnalyseParams) * len(aps))(*[C.pointer(a) for a in aps])
    outs = (C.c_void_p * len(aps))(*[t.data_ptr() for t in dvec])
    def step():
        mb._check(L.mvb200_super_batch(ctx.handle, C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), N))
        ctx.synchronize(); t1 = time.time()
        mb._check(L.mvb200_analyse_multi_batch(ctx.handle, apa, len(aps), C.byref(sp), dsup.data_ptr(), 0, N, outs, vs))
        ctx.synchronize(); t2 = time.time()
        mb._check(L.mvb200_degrain_batch(ctx.handle, C.byref(dp), C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), N, 0, N, outs, vs, ddst.data_ptr()))
        ctx.synchronize(); t3 = time.time()
        return t1, t2, t3
    step()
    t0 = time.time(); t1, t2, t3 = step()
    print(f"{W}x{H} bits={bits} N={N} r={radius} blk={blk} ov={ov} pel={pel}: super {1e3*(t1-t0):.1f} ms, analyse {1e3*(t2-t1):.1f} ms, degrain {1e3*(t3-t2):.1f} ms, total fps {N/(t3-t0):.2f}", flush=True)
    if check:
        import oracle
        n = N // 2
        sup = dsup[n].cpu().numpy()
        def planes_of(buf, lay, bits):
            out = []
            for p in range(lay.nplanes):
                a = buf[lay.offset[p]: lay.offset[p] + lay.pitch[p] * lay.plane_h[p]].reshape(lay.plane_h[p], lay.pitch[p])[:, : lay.plane_w[p] * lay.bytes_per_sample]
                a = np.ascontiguousarray(a)
                out.append(a.view(np.uint16) if lay.bytes_per_sample == 2 else a)
            return out
        osups = {m: oracle.super_frame(sp, frames[m]) for m in range(max(0, n - radius), min(N, n + radius + 1))}
        g = planes_of(sup, sl, bits)
        print("  super match:", all(np.array_equal(a, b) for a, b in zip(g, osups[n])))
        ovs = []
        for i, ap in enumerate(aps):
            d = ap.ad.nDeltaFrame; nref = n + d if ap.ad.isBackward else n - d
            t = time.time()
            ov_ = oracle.analyse_frame(ap, sp, osups[n], osups[nref] if 0 <= nref < N else None)
            gv = dvec[i][n].cpu().numpy()[: ap.vectors_size]
            ovs.append(ov_)
            print(f"  analyse[{i}] match:", np.array_equal(gv, ov_), f"(oracle {time.time()-t:.2f}s)")
        refs = []
        for ap in aps:
            d = ap.ad.nDeltaFrame; nref = n + d if ap.ad.isBackward else n - d
            refs.append(osups[nref] if 0 <= nref < N else None)
        oo = oracle.degrain_frame(dp, sp, frames[n], refs, ovs)
        gd = planes_of(ddst[n].cpu().numpy(), cl, bits)
        print("  degrain match:", all(np.array_equal(a, b) for a, b in zip(gd, oo)))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "configs":  # resident-clip throughput of BASELINE.json's other configurations (not a bench line)
        run(1920, 1080, 3552, 8, 1, 8, 4, 2, check=False, distinct=64)     # configs[1]: Degrain1 blk 8 ov 4 pel 2, 1080p P8
        run(1920, 1080, 222, 16, 3, 16, 8, 2, check=False, distinct=64)    # configs[2]: Degrain3 blk 16 ov 8 pel 2, 1080p P16
        run(3840, 2160, 64, 16, 3, 16, 8, 2, check=False, distinct=16)     # configs[4]: one GPU's 64-frame shard of the 4K P16 batch
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "prof":  # one 1080p P8 Degrain3 pass for ncu
        run(1920, 1080, int(sys.argv[2]) if len(sys.argv) > 2 else 64, 8, 3, 16, 8, 2, check=False)
        sys.exit(0)
    sys.path.insert(0, ROOT); import __graft_entry__ as g
    g.smoke()
    run(640, 360, 8, 8, 1, 8, 0, 1)
    run(640, 360, 8, 8, 1, 8, 4, 2)
    run(1920, 1080, 8, 8, 1, 8, 4, 2)
    run(1920, 1080, 16, 16, 3, 16, 8, 2)
    run(1920, 1080, 64, 8, 3, 16, 8, 2, check=False)
