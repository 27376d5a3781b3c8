"""Reproducer / regression stress for the table-upload race (csrc/engine.h upload_table): the per-frame mv.Compensate and
mv.Degrain of two configurations with different overlap windows, ALTERNATING, so that every call takes the "first use of this
configuration" path (rebuild the windows on the host, upload, launch) -- compared with the reference's golden hashes.
    python tools/stress_tables.py [reps]                                  (the product library: expect 0 mismatches)
    MVB200_LIB=libmvtools_b200_synctab.so python tools/stress_tables.py   (built with -DMVB_SYNC_TABLE_UPLOAD: the old upload)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("vapoursynth-mvtools_b200", "oracle", os.path.join("host", "minivs"), "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import pipeline
import mvtools_b200 as mb

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
ctx = mb.Context(0)
prep = {}
for case in ("blk128x64_gray_p8", "blk64_420_ov32_r2"):
    frames = pipeline.make_frames(pipeline.CASES[case][0])
    N = len(frames)
    sp, aps, dp = pipeline.params_for(case)
    rp, cp = pipeline.extra_params_for(case)
    sups = [ctx.super_frame(sp, f) for f in frames]
    nref = [pipeline.nref_of(aps[0], n, N) for n in range(N)]
    vecs = [ctx.analyse_frame(aps[0], sp, sups[n], sups[nref[n]] if nref[n] is not None else None) for n in range(N)]
    prep[case] = (sp, cp, sups, nref, vecs)
bad = {c: 0 for c in prep}
for it in range(reps):
    for case, (sp, cp, sups, nref, rvecs) in prep.items():
        n = 0
        got = ctx.compensate_frame(cp, sp, sups[n], sups[nref[n]] if nref[n] is not None else None, rvecs[n])
        if [pipeline.sha(p) for p in got] != GOLD[case]["compensate"][n]:
            bad[case] += 1
print("library %s: %d alternating calls per case, mismatches %s" % (os.environ.get("MVB200_LIB", "libmvtools_b200.so"), reps, bad))
