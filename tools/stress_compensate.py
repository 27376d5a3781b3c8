"""Hunts an intermittent mismatch seen ONCE in the per-frame mv.Compensate GPU test (late case blk128x64_gray_p8, frame 0): runs the
plugin's windows first (several engine contexts with work in flight), then the per-frame Compensate of that case many times and
compares every result with the reference's golden hash; prints where a differing result differs."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("vapoursynth-mvtools_b200", "oracle", os.path.join("host", "minivs"), "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import bench, minivs, pipeline
import mvtools_b200 as mb

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
ctx = mb.Context(0)
# background: the 1080p script through the plugin, left with windows in flight
cfg = bench.CONFIGS["headline"]
core = minivs.Core()
mv = core.load(mb.PLUGIN_PATH)
out, keep = bench.build_script(mv, core, cfg, bench.make_frames(cfg, 1500))
out.set_cache(0)
out.render(0, 300, 8)
bad = 0
for case in ("blk128x64_gray_p8", "blk64_420_ov32_r2", "p8_440_blk8_ov4_r2"):
    frames = pipeline.make_frames(pipeline.CASES[case][0])
    N = len(frames)
    sp, aps, dp = pipeline.params_for(case)
    rp, cp = pipeline.extra_params_for(case)
    sups = [ctx.super_frame(sp, f) for f in frames]
    vecs = [ctx.analyse_frame(aps[0], sp, sups[n], sups[pipeline.nref_of(aps[0], n, N)] if pipeline.nref_of(aps[0], n, N) is not None else None) for n in range(N)]
    first = None
    for it in range(reps):
        if it == reps // 2:
            out.render(300, 700, 8)  # more plugin work in between
        for n in range(N):
            r = pipeline.nref_of(aps[0], n, N)
            got = ctx.compensate_frame(cp, sp, sups[n], sups[r] if r is not None else None, vecs[n])
            h = [pipeline.sha(p) for p in got]
            if h != GOLD[case]["compensate"][n]:
                bad += 1
                ref = first[n] if first else None
                where = [np.argwhere(a != b)[:6].tolist() for a, b in zip(got, ref)] if ref else "?"
                print("MISMATCH", case, "iteration", it, "frame", n, "differs at", where, flush=True)
            elif first is None or len(first) <= n:
                first = (first or []) + [got]
print("stress done:", bad, "mismatches")
