"""One Super + 2r x Analyse + DegrainN pass at a BASELINE configuration for ncu (not a bench):
    python tools/gpu_prof2.py [frames] [bits] [blk] [ov] [radius] [WxH]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vapoursynth-mvtools_b200"))
import numpy as np, torch
import clipgen, mvtools_b200 as mb
F = int(sys.argv[1]) if len(sys.argv) > 1 else 32
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 8
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 16
ov = int(sys.argv[4]) if len(sys.argv) > 4 else 8
radius = int(sys.argv[5]) if len(sys.argv) > 5 else 3
W, H = (int(x) for x in (sys.argv[6] if len(sys.argv) > 6 else "1920x1080").split("x"))
L = mb.lib()
DISTINCT = min(F, 64)
base = clipgen.make_clip(W, H, DISTINCT, bits=bits, seed=0)
frames = [base[i % DISTINCT] for i in range(F)]
sp = mb.super_params(W, H, bits, 1, 1, pel=2)
cl, sl = mb.clip_layout(sp), mb.super_layout(sp)
ctx = mb.Context(0)
host = np.zeros((F, cl.frame_bytes), dtype=np.uint8)
for n, fr in enumerate(frames):
    for p, a in enumerate(fr):
        rows = host[n, cl.offset[p]: cl.offset[p] + cl.pitch[p] * cl.plane_h[p]].reshape(cl.plane_h[p], cl.pitch[p])
        rows[:, : a.shape[1] * a.itemsize] = a.view(np.uint8).reshape(a.shape[0], -1)
dsrc = torch.from_numpy(host).cuda()
dsup = torch.empty((F, sl.frame_bytes), dtype=torch.uint8, device="cuda")
ddst = torch.empty((F, cl.frame_bytes), dtype=torch.uint8, device="cuda")
aps = [mb.analyse_params(sp, F, isb=isb, delta=d, blksize=blk, overlap=ov) for d in range(1, radius + 1) for isb in (1, 0)]
dp = mb.degrain_params(sp, [a.ad for a in aps], W, H, radius)
vs = (aps[0].vectors_size + 255) // 256 * 256
dvec = [torch.empty((F, vs), dtype=torch.uint8, device="cuda") for _ in aps]
apa = (C.POINTER(mb.AnalyseParams) * len(aps))(*[C.pointer(a) for a in aps])
outs = (C.c_void_p * len(aps))(*[t.data_ptr() for t in dvec])
mb._check(L.mvb200_super_batch(ctx.handle, C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), F))
mb._check(L.mvb200_analyse_multi_batch(ctx.handle, apa, len(aps), C.byref(sp), dsup.data_ptr(), 0, F, outs, vs))
mb._check(L.mvb200_degrain_batch(ctx.handle, C.byref(dp), C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), F, 0, F, outs, vs, ddst.data_ptr()))
ctx.synchronize()
print("done", F, "frames,", len(aps) * F, "chains")
