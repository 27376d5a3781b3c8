"""Shared helpers for the consumers of the vector clips (mv.SCDetection, mv.Mask, mv.Finest,
mv.Flow): run them on the seeded clips of pipeline.CASES through
 (a) a VapourSynth plugin (the compiled reference by default) and the mini host,
 (b) per-frame stage functions (the CPU oracle, or the CUDA C ABI),
and hash every output the same way, so results can be compared or pinned as goldens
(tests/golden/golden_flow.json, made by tests/golden/make_golden_flow.py)."""
import functools

import numpy as np

import mvtools_b200 as mb
import pipeline
from pipeline import CASES, sha

# case -> stages.  `vec`: which vector clip feeds the filter (0 = backward delta 1, 1 = forward delta 1).
# Mask needs an 8-bit clip (MVMask.c:318); the reference's mv.Finest overruns subsampled chroma planes at pel = 1
# (MVFinest.c:75-77 copies the luma size for every plane), so `finest` is only pinned for pel > 1 or 4:4:4 / gray.
FLOW = {
    "blk8_ov4_pel2_p8": dict(
        mask=[dict(kind=0), dict(kind=1, gamma=2.0, time=30.0), dict(kind=2, ml=50.0), dict(kind=2, vec=1, gamma=0.7, time=60.0),
              dict(kind=3), dict(kind=4, ml=20.0, vec=1), dict(kind=5, ysc=77), dict(kind=0, gamma=1.7, ml=30.0, thscd1=200, thscd2=40, ysc=200)],
        flow=[dict(), dict(time=50.0, vec=1), dict(mode=1, time=70.0), dict(mode=1, vec=1, thscd1=250, thscd2=60), dict(fields=1, tff=1, time=80.0), dict(fields=1, tff=0, vec=1, mode=1)],
        scd=[dict(), dict(vec=1, thscd1=150, thscd2=30)], finest=True,
        inter=[dict(), dict(time=25.0, ml=40.0), dict(time=80.0, blend=0, thscd1=250, thscd2=60), dict(time=0.0), dict(time=100.0)],
        blur=[dict(), dict(blur=120.0, prec=2), dict(blur=200.0, thscd1=250, thscd2=60)],
        bfps=[dict(num=50, den=1), dict(num=60, den=1, mode=0), dict(num=60, den=1, mode=1), dict(num=60, den=1, mode=2), dict(num=40, den=1, mode=4, ml=30.0), dict(num=60, den=1, mode=5),
              dict(num=60, den=1, mode=6, ml=50.0), dict(num=40, den=1, mode=7), dict(num=60, den=1, mode=8, blend=0, thscd1=250, thscd2=60)],
        fps=[dict(num=50, den=1), dict(num=60, den=1, mask=0), dict(num=30000, den=1001, mask=1, ml=50.0), dict(num=50, den=1, mask=0), dict(num=40, den=1, mask=2, blend=0, thscd1=250, thscd2=60)]),
    "blk16_ov8_pel2_p16_r3": dict(flow=[dict(time=33.0), dict(mode=1, vec=1)], scd=[dict()], finest=True, inter=[dict(), dict(time=70.0, ml=20.0)], fps=[dict(num=50, den=1, mask=0), dict(num=75, den=2, mask=1)], blur=[dict(blur=80.0)], bfps=[dict(num=60, den=1), dict(num=50, den=1, mode=7, ml=40.0), dict(num=40, den=1, mode=2)]),
    "p10_444_blk8_pel4_bicubic": dict(flow=[dict(time=80.0), dict(mode=1, time=45.0, vec=1)], finest=True, scd=[dict(thscd2=10)], inter=[dict(time=40.0)], fps=[dict(num=60, den=1)], blur=[dict(blur=150.0, prec=3)], bfps=[dict(num=60, den=1, mode=4), dict(num=60, den=1, mode=1)]),
    "gray_blk4_pel1_rf0": dict(mask=[dict(kind=0), dict(kind=2, time=50.0), dict(kind=5)], flow=[dict(), dict(mode=1, vec=1, time=25.0)], finest=True, scd=[dict()], inter=[dict(time=60.0, ml=10.0)], blur=[dict()], bfps=[dict(num=60, den=1), dict(num=60, den=1, mode=8)]),
    "c1_640x360_p8_blk8_pel1": dict(mask=[dict(kind=1), dict(kind=2, vec=1)], flow=[dict(time=60.0), dict(mode=1)], scd=[dict()], inter=[dict(), dict(time=33.0, blend=0)], bfps=[dict(num=40, den=1), dict(num=40, den=1, mode=6), dict(num=40, den=1, mode=4)]),
    "p8_422_blk8x4_sharp0_rf1": dict(mask=[dict(kind=0, ml=10.0), dict(kind=1, time=0.0), dict(kind=2, vec=1, ml=10.0), dict(kind=5)], flow=[dict(vec=1), dict(mode=1, time=100.0)], finest=True,
                                     scd=[dict(vec=1)], inter=[dict(time=45.0)], bfps=[dict(num=60, den=1, mode=3), dict(num=60, den=1, mode=2)]),
    "noise_badblocks_blk8": dict(mask=[dict(kind=0, ysc=9), dict(kind=1, thscd1=16000, thscd2=255), dict(kind=2, thscd1=16000, thscd2=255, ml=400.0), dict(kind=5, thscd1=16000, thscd2=255, ml=3.0)],
                                 flow=[dict(), dict(thscd1=16000, thscd2=255, time=90.0), dict(mode=1, thscd1=16000, thscd2=255, vec=1)], scd=[dict(), dict(thscd1=16000, thscd2=255)],
                                 inter=[dict(), dict(blend=0), dict(thscd1=16000, thscd2=255, ml=300.0, time=20.0)], fps=[dict(num=60, den=1), dict(num=60, den=1, blend=0), dict(num=50, den=1, thscd1=16000, thscd2=255, mask=1)], blur=[dict(), dict(thscd1=16000, thscd2=255, blur=100.0)], bfps=[dict(num=60, den=1), dict(num=60, den=1, thscd1=16000, thscd2=255, mode=7), dict(num=60, den=1, blend=0)]),
    "blk4_420_chroma2x2_pel2": dict(mask=[dict(kind=2, gamma=2.0)], flow=[dict(time=12.0), dict(mode=1)], finest=True),
    # (no mask here: 256-sample rows equal the stride, and the reference's bottom-strip copy, MVMask.c:162-163, is an
    # overlapping memcpy then -- undefined behaviour; the oracle and the engine replicate the last covered row, which is
    # what the reference's row loop does for every other width)
    "divide2_blk16_esa_neg_badrange": dict(flow=[dict(thscd1=16000, thscd2=255), dict(mode=1, thscd1=16000, thscd2=255, vec=1, time=40.0)], scd=[dict(thscd1=16000, thscd2=255)]),
    # pipeline.LATE cases (large and flat blocks): GPU tests in test_zz_gpu_pending.py
    "blk64_420_ov32_r2": dict(mask=[dict(kind=0), dict(kind=1, time=40.0), dict(kind=2, vec=1, ml=60.0), dict(kind=5)], flow=[dict(time=70.0), dict(mode=1, vec=1)], scd=[dict()], finest=True,
                              inter=[dict(), dict(time=30.0, ml=50.0, blend=0)], fps=[dict(num=50, den=1), dict(num=40, den=1, mask=1)], blur=[dict(blur=90.0)], bfps=[dict(num=50, den=1), dict(num=50, den=1, mode=4), dict(num=40, den=1, mode=7, ml=60.0)]),
    "blk16x8_ov8x4_p10": dict(flow=[dict(), dict(mode=1, vec=1, time=35.0)], scd=[dict(vec=1)], finest=True, inter=[dict(time=60.0)], fps=[dict(num=60, den=1, mask=0)], blur=[dict()], bfps=[dict(num=60, den=1), dict(num=60, den=1, mode=2)]),
    # (no flow / inter / fps with 128-wide blocks: the reference corrupts its heap there -- glibc aborts with "double free or
    # corruption" -- whenever the vector field is narrower than 8 blocks, which its SIMD SimpleResize rows assume)
    "blk128x64_gray_p8": dict(mask=[dict(kind=0), dict(kind=2, time=70.0), dict(kind=3), dict(kind=4, vec=1)], scd=[dict()], finest=True, blur=[dict(blur=60.0, prec=2)], bfps=[dict(num=45, den=1), dict(num=45, den=1, mode=6)]),
    "blk64x32_ov_vertical_p8": dict(mask=[dict(kind=1), dict(kind=2, ml=80.0)], flow=[dict(), dict(mode=1, time=20.0, vec=1)], scd=[dict()], inter=[dict(time=45.0, ml=30.0)], fps=[dict(num=45, den=1)]),
    "p8_440_blk8_ov4_r2": dict(mask=[dict(kind=0), dict(kind=2, vec=1), dict(kind=5)], flow=[dict(time=55.0), dict(mode=1, vec=1)], scd=[dict()], finest=True, inter=[dict(), dict(time=30.0, ml=60.0)],
                               fps=[dict(num=50, den=1), dict(num=40, den=1, mask=1)], blur=[dict(blur=70.0)], bfps=[dict(num=50, den=1), dict(num=50, den=1, mode=4)]),
    "p8_420_blk8_pel4": dict(flow=[dict(time=40.0), dict(mode=1, vec=1)], finest=True, scd=[dict()], inter=[dict(time=30.0)], blur=[dict()], bfps=[dict(num=50, den=1)]),
    "p8_420_blk16_ov4_uncovered": dict(mask=[dict(kind=0), dict(kind=1, ml=30.0), dict(kind=2, vec=1), dict(kind=5)], flow=[dict(), dict(mode=1, vec=1, time=66.0)], scd=[dict()], finest=True, inter=[dict(), dict(time=10.0, ml=50.0)], fps=[dict(num=35, den=1, mask=2), dict(num=20, den=1, mask=0)], blur=[dict(blur=66.0)], bfps=[dict(num=35, den=1), dict(num=35, den=1, mode=1), dict(num=35, den=1, mode=5)]),
}


def _split(kw):
    kw = dict(kw)
    return kw.pop("vec", 0), kw


@functools.lru_cache(maxsize=None)
def inputs(case):
    """(frames, super frames, [vectors of clip 0 per frame, vectors of clip 1 per frame]) computed by the CPU oracle
    (itself pinned against the reference for these stages); inputs of the per-frame stage functions."""
    import oracle
    frames = pipeline.make_frames(CASES[case][0])
    N = len(frames)
    sp, aps, dp = pipeline.params_for(case)
    if case in pipeline.PELCLIP:
        raise ValueError("no flow stages for pelclip cases")
    sups = [oracle.super_frame(sp, f) for f in frames]
    vecs = []
    for ap in aps[:2]:
        row = []
        for n in range(N):
            r = pipeline.nref_of(ap, n, N)
            row.append(oracle.analyse_frame(ap, sp, sups[n], sups[r] if r is not None else None))
        vecs.append(row)
    return frames, sups, vecs


def flow_params_for(case):
    """-> sp, aps, {stage: [params per variant]}"""
    ck = CASES[case][0]
    N = ck["nframes"]
    sp, aps, dp = pipeline.params_for(case)
    vds = [(ap.ad_divided if ap.divide else ap.ad) for ap in aps[:2]]
    out = {"mask": [], "flow": [], "scd": [], "inter": [], "fps": [], "blur": [], "bfps": []}
    for kw in FLOW[case].get("bfps", []):
        out["bfps"].append(mb.blockfps_params(sp, vds[0], vds[1], N, ck["width"], ck["height"], **kw))
    for kw in FLOW[case].get("blur", []):
        out["blur"].append(mb.flowblur_params(sp, vds[0], vds[1], N, **kw))
    for kw in FLOW[case].get("fps", []):
        out["fps"].append(mb.flowfps_params(sp, vds[0], vds[1], N, ck["width"], ck["height"], **kw))
    for kw in FLOW[case].get("inter", []):
        out["inter"].append(mb.flowinter_params(sp, vds[0], vds[1], N, **kw))
    for kw in FLOW[case].get("mask", []):
        v, kw = _split(kw)
        out["mask"].append((v, mb.mask_params(vds[v], clip_bits=ck.get("bits", 8), clip_ssw=ck.get("ssw", 1), clip_ssh=ck.get("ssh", 1), clip_gray=ck.get("gray", False), **kw)))
    for kw in FLOW[case].get("flow", []):
        v, kw = _split(kw)
        out["flow"].append((v, mb.flow_params(sp, vds[v], N, **kw)))
    for kw in FLOW[case].get("scd", []):
        v, kw = _split(kw)
        out["scd"].append((v, mb.scdetection_params(vds[v], **kw)))
    return sp, aps, out


def run_engine_inter(case, inter_fn, blend_fn):
    """mv.FlowInter frame by frame (MVFlowInter.c:83-448): stage functions on the oracle-made inputs -> hashes"""
    frames, sups, vecs = inputs(case)
    N = len(frames)
    sp, aps, P = flow_params_for(case)
    out = []
    for p in P["inter"]:
        off = p.vd.nDeltaFrame
        rows = []
        for n in range(N):
            d = None
            if n + off < N:  # vec_b = mvbw(n), vec_f = mvfw(n + off), extra pair: mvbw(n + off), mvfw(n)
                d = inter_fn(p, sp, p.time256, sups[n], sups[n + off], vecs[0][n], vecs[1][n + off], vecs[0][n + off], vecs[1][n])
            if d is None:  # :412-447
                d = blend_fn(sp, p.time256, frames[n], frames[min(n + off, N - 1)]) if p.blend else frames[n]
            rows.append([sha(q) for q in d])
        out.append(rows)
    return out


def run_engine_fps(case, inter_fn, blend_fn):
    """mv.FlowFPS output frame by output frame (MVFlowFPS.c:86-470) -> hashes"""
    frames, sups, vecs = inputs(case)
    N = len(frames)
    sp, aps, P = flow_params_for(case)
    out = []
    for p in P["fps"]:
        q, off = p.inter, p.inter.vd.nDeltaFrame
        rows = []
        for n in range(p.num_frames_out):
            nleft, time256 = mb.flowfps_position(p, n)
            nright = nleft + off
            if time256 == 0:
                d = frames[min(nleft, N - 1)]
            elif time256 == 256:
                d = frames[min(nright, N - 1)]
            else:
                d = None
                if nleft < N and nright < N:
                    extra = p.maskmode == 2
                    d = inter_fn(q, sp, time256, sups[nleft], sups[nright], vecs[0][nleft], vecs[1][nright], vecs[0][nright] if extra else None, vecs[1][nleft] if extra else None)
                if d is None:  # :470-520
                    src = frames[min(nleft, N - 1)]
                    d = blend_fn(sp, time256, src, frames[min(nright, N - 1)]) if q.blend else src
            rows.append([sha(x) for x in d])
        out.append(rows)
    return out


def run_engine_blur(case, blur_fn):
    """mv.FlowBlur frame by frame (MVFlowBlur.c:144-337): mvbw of n - delta and mvfw of n + delta -> hashes"""
    frames, sups, vecs = inputs(case)
    N = len(frames)
    sp, aps, P = flow_params_for(case)
    out = []
    for p in P["blur"]:
        off = p.vd.nDeltaFrame
        rows = []
        for n in range(N):
            d = blur_fn(p, sp, sups[n], vecs[0][n - off], vecs[1][n + off]) if n - off >= 0 and n + off < N else None
            rows.append([sha(q) for q in (d if d is not None else frames[n])])
        out.append(rows)
    return out


def run_engine_bfps(case, bfps_fn, blend_fn):
    """mv.BlockFPS output frame by output frame (MVBlockFPS.c:240-680) -> hashes"""
    frames, sups, vecs = inputs(case)
    N = len(frames)
    sp, aps, P = flow_params_for(case)
    nplanes_used = 3 if (sp.modeyuv & 6) else 1
    out = []
    for p in P["bfps"]:
        off = p.vd.nDeltaFrame
        rows = []
        for n in range(p.num_frames_out):
            nleft, time256 = mb.blockfps_position(p, n)
            nright = nleft + off
            if time256 == 0:
                d = frames[min(nleft, N - 1)]
            elif time256 == 256:
                d = frames[min(nright, N - 1)]
            else:
                d = bfps_fn(p, sp, time256, sups[nleft], sups[nright], vecs[0][nleft], vecs[1][nright]) if nleft < N and nright < N else None
                if d is None:  # :640-670
                    src = frames[min(nleft, N - 1)]
                    d = blend_fn(sp, time256, src, frames[min(nright, N - 1)]) if p.blend else src
            rows.append([sha(x) for x in d[:nplanes_used]])
        out.append(rows)
    return out


def run_engine_flow(case, scd_fn, mask_fn, finest_fn, flow_fn):
    """Per-frame stage functions (oracle.* or Context.*) on the oracle-made inputs -> {stage: hashes}"""
    frames, sups, vecs = inputs(case)
    N = len(frames)
    sp, aps, P = flow_params_for(case)
    out = {"scd": [], "mask": [], "flow": [], "finest": []}
    for v, scp in P["scd"]:
        out["scd"].append([int(scd_fn(scp, vecs[v][n])) for n in range(N)])
    for v, mp in P["mask"]:
        out["mask"].append([[sha(p) for p in mask_fn(mp, frames[n][0], vecs[v][n])] for n in range(N)])
    for v, fp in P["flow"]:
        rows = []
        for n in range(N):
            r = pipeline.nref_of(aps[v], n, N)
            fs = 0
            if r is not None and fp.fields and fp.vd.nPel > 1 and (r - n) % 2:  # MVFlow.cpp:268-299 with an explicit tff
                src_top, ref_top = fp.tff ^ (n % 2), fp.tff ^ (r % 2)
                fs = fp.vd.nPel // 2 if (src_top and not ref_top) else (-(fp.vd.nPel // 2) if (ref_top and not src_top) else 0)
            d = flow_fn(fp, sp, sups[r] if r is not None else None, vecs[v][n], fs)
            rows.append([sha(p) for p in (d if d is not None else frames[n])])  # not usable: the source frame (MVFlow.cpp:362-366)
        out["flow"].append(rows)
    if FLOW[case].get("finest"):
        out["finest"] = [[sha(p) for p in finest_fn(sp, sups[n])] for n in range(N)]
    return out


def run_reference_flow(case, plugin_path=None, only=None):
    """The same stages through a VapourSynth plugin (default: the compiled reference) and the mini host.
    only: restrict to these stage names."""
    import minivs
    ck, skw, akw, radius, dkw = CASES[case]
    frames = pipeline.make_frames(ck)
    N = len(frames)
    core = minivs.Core()
    mv = core.load(plugin_path or minivs.REF_PLUGIN)
    clip = core.source(frames, bits=ck.get("bits", 8), ssw=ck.get("ssw", 1), ssh=ck.get("ssh", 1))
    sup = mv.Super(clip, **skw)
    vn = [mv.Analyse(sup, isb=isb, delta=1, **pipeline._fix(akw)) for isb in (1, 0)]
    nplanes = 1 if ck.get("gray") else 3
    out = {"scd": [], "mask": [], "flow": [], "finest": []}
    want = lambda stage: FLOW[case].get(stage, []) if (only is None or stage in only) else []
    for kw in want("scd"):
        v, kw = _split(kw)
        node = mv.SCDetection(clip, vn[v], **kw)
        key = "_SceneChangeNext" if v == 0 else "_SceneChangePrev"  # isBackward -> Next (MVSCDetection.c:61-63)
        out["scd"].append([int(node.get_frame(n).prop_int(key)) for n in range(N)])
    for kw in want("mask"):
        v, kw = _split(kw)
        kw = {k: (float(x) if k in ("ml", "gamma", "time") else x) for k, x in kw.items()}
        node = mv.Mask(clip, vn[v], **kw)
        out["mask"].append([[sha(p) for p in node.get_frame(n).planes()] for n in range(N)])
    for kw in want("flow"):
        v, kw = _split(kw)
        kw = {k: (float(x) if k == "time" else x) for k, x in kw.items()}
        node = mv.Flow(clip, sup, vn[v], **kw)
        out["flow"].append([[sha(p) for p in node.get_frame(n).planes()[:nplanes]] for n in range(N)])
    if want("finest"):
        node = mv.Finest(sup)
        out["finest"] = [[sha(p) for p in node.get_frame(n).planes()[:nplanes]] for n in range(N)]
    out["inter"] = []
    for kw in want("inter"):
        kw = {k: (float(x) if k in ("time", "ml") else x) for k, x in kw.items()}
        node = mv.FlowInter(clip, sup, vn[0], vn[1], **kw)
        out["inter"].append([[sha(p) for p in node.get_frame(n).planes()[:nplanes]] for n in range(N)])
    out["blur"] = []
    for kw in want("blur"):
        kw = {k: (float(x) if k == "blur" else x) for k, x in kw.items()}
        node = mv.FlowBlur(clip, sup, vn[0], vn[1], **kw)
        out["blur"].append([[sha(p) for p in node.get_frame(n).planes()[:nplanes]] for n in range(N)])
    out["bfps"] = []
    for kw in want("bfps"):
        kw = {k: (float(x) if k == "ml" else x) for k, x in kw.items()}
        node = mv.BlockFPS(clip, sup, vn[0], vn[1], **kw)
        out["bfps"].append([[sha(p) for p in node.get_frame(n).planes()[:nplanes]] for n in range(node.num_frames)])
    out["fps"] = []
    for kw in want("fps"):
        kw = {k: (float(x) if k == "ml" else x) for k, x in kw.items()}
        node = mv.FlowFPS(clip, sup, vn[0], vn[1], **kw)
        out["fps"].append([[sha(p) for p in node.get_frame(n).planes()[:nplanes]] for n in range(node.num_frames)])
    return out
