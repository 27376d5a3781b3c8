#!/usr/bin/env python
"""How long one mvb200_analyse_clip_host call takes for a window of configs[0] (640x360 P8, pel 1, blk 8), alone and with two
contexts side by side (what the plugin's fused Analyse path does)."""
import ctypes as C, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vapoursynth-mvtools_b200")]
import torch
import mvtools_b200 as mb
import bench

L = mb.lib()
L.mvb200_analyse_clip_host.argtypes = [C.c_void_p, C.POINTER(mb.SuperParams), C.POINTER(mb.AnalyseParams), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
cfg = bench.CONFIGS["c1"]
W = int(sys.argv[1]) if len(sys.argv) > 1 else 768
n = W + 1
sp = mb.super_params(cfg["w"], cfg["h"], cfg["bits"], 1, 1, pel=cfg["pel"])
ap = mb.analyse_params(sp, 100000, blksize=cfg["blk"], overlap=cfg["ov"], search=cfg.get("search", 4))
cl, sl = mb.clip_layout(sp), mb.super_layout(sp)
vs = (ap.vectors_size + 255) // 256 * 256
frames = bench.make_frames(cfg, n, first=0)


def buffers():
    host = torch.zeros((n, cl.frame_bytes), dtype=torch.uint8).pin_memory()
    bench.fill_host_clip(host.numpy(), frames, cl)
    return host, torch.zeros((W, sl.frame_bytes), dtype=torch.uint8).pin_memory(), torch.zeros((W, vs), dtype=torch.uint8).pin_memory()


def call(ctx, b, first):
    t = time.perf_counter()
    mb._check(L.mvb200_analyse_clip_host(ctx.handle, C.byref(sp), C.byref(ap), b[0].data_ptr(), first, n, first + 1, W, b[1].data_ptr(), b[2].data_ptr(), vs))
    return (time.perf_counter() - t) * 1e3


ctxs = [mb.Context(0), mb.Context(0)]
bufs = [buffers(), buffers()]
print("window %d frames: H2D %.0f MB, D2H %.0f MB" % (W, n * cl.frame_bytes / 1e6, W * (sl.frame_bytes + vs) / 1e6))
print("one context, back to back (ms):", ["%.1f" % call(ctxs[0], bufs[0], 1000) for _ in range(6)])
res = [[], []]


def worker(i):
    for _ in range(6):
        res[i].append(call(ctxs[i], bufs[i], 1000))


t0 = time.perf_counter()
th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
[t.start() for t in th]
[t.join() for t in th]
dt = time.perf_counter() - t0
print("two contexts side by side (ms):", [["%.1f" % x for x in r] for r in res], "-> %.0f frames/s" % (12 * W / dt))
# stage times of one call on one context (CUDA events on the context's stream are not reachable from here: host clocks around
# the pieces through the batch entry points)
st = torch.cuda.ExternalStream(ctxs[0].stream())
dsrc = bufs[0][0].to("cuda")
dsup = torch.empty((n, sl.frame_bytes), dtype=torch.uint8, device="cuda")
dvec = torch.empty((W, vs), dtype=torch.uint8, device="cuda")
a2 = mb.analyse_params(sp, n, blksize=cfg["blk"], overlap=cfg["ov"], search=cfg.get("search", 4))
for name, fn in (("h2d", lambda: dsrc.copy_(bufs[0][0], non_blocking=True)),
                 ("super", lambda: mb._check(L.mvb200_super_batch(ctxs[0].handle, C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), n))),
                 ("analyse", lambda: mb._check(L.mvb200_analyse_batch(ctxs[0].handle, C.byref(a2), C.byref(sp), dsup.data_ptr(), 1, W, dvec.data_ptr(), vs))),
                 ("d2h super", lambda: bufs[0][1].copy_(dsup[1:], non_blocking=True))):
    ts = []
    for _ in range(4):
        torch.cuda.synchronize(); ctxs[0].synchronize()
        t = time.perf_counter()
        with torch.cuda.stream(st):
            fn()
        ctxs[0].synchronize(); torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    print("%-10s %s ms" % (name, ["%.1f" % x for x in ts]))
