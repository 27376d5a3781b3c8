"""GPU parity of the BENCH workload itself, at full size: bench.py's default configuration (1920x1080 YUV420P8, Super pel 2,
6 x Analyse blk 16 / ov 8, Degrain3, the 64-frame cyclic clip of clipgen seed 0) against sha256 digests of the compiled
reference's output (tests/golden/bench_digest.json, made by tests/golden/make_golden_bench.py from oracle/_ref).  The frames
span the scene cut at 32 and the 63 -> 64 wrap of the cycle.  Three routes: the clip-level C ABI with host buffers, the
device-resident batch calls bench.py times, and the drop-in plugin in its default configuration (look-ahead windows)."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import bench
import mvtools_b200 as mb
import sharding

pytestmark = pytest.mark.gpu

N = 80  # frames 0 .. 79 of the cyclic clip


def _planes_digest(buf, cl):
    h = hashlib.sha256()
    for p in range(cl.nplanes):
        rows = buf[cl.offset[p]: cl.offset[p] + cl.pitch[p] * cl.plane_h[p]].reshape(cl.plane_h[p], cl.pitch[p])
        h.update(np.ascontiguousarray(rows[:, : cl.plane_w[p] * cl.bytes_per_sample]).tobytes())
    return h.hexdigest()


def _setup(first, count, total):
    cfg = bench.CONFIGS["headline"]
    plan = sharding.local_plan(total, 1, 0, reach=bench.REACH) if first is None else None
    lo = 0 if first is None else max(0, first - bench.REACH)
    hi = total if first is None else min(total, first + count + bench.REACH)
    F = hi - lo
    sp = mb.super_params(cfg["w"], cfg["h"], cfg["bits"], 1, 1, pel=cfg["pel"])
    aps = [mb.analyse_params(sp, F, isb=isb, delta=d, blksize=cfg["blk"], overlap=cfg["ov"]) for d in range(1, cfg["radius"] + 1) for isb in (1, 0)]
    dp = mb.degrain_params(sp, [a.ad for a in aps], cfg["w"], cfg["h"], cfg["radius"])
    cl = mb.clip_layout(sp)
    host = np.zeros((F, cl.frame_bytes), dtype=np.uint8)
    bench.fill_host_clip(host, bench.make_frames(cfg, F, first=lo), cl)
    return cfg, sp, aps, dp, cl, host, lo, F, plan


def test_bench_workload_clip_call_matches_reference_digests(gpu_ctx):
    gold = bench.load_golden("headline")
    assert gold is not None
    cfg, sp, aps, dp, cl, host, lo, F, _ = _setup(None, None, N)
    L = mb.lib()
    apa = (C.POINTER(mb.AnalyseParams) * len(aps))(*[C.pointer(a) for a in aps])
    out = np.zeros((F, cl.frame_bytes), dtype=np.uint8)
    mb._check(L.mvb200_degrain_clip_host_async(gpu_ctx.handle, C.byref(sp), apa, len(aps), C.byref(dp), host.ctypes.data, out.ctypes.data, F, 0, F))
    gpu_ctx.synchronize()
    checked = 0
    for g in range(N):
        want = bench.expected_digest(gold, g, N)
        if want is not None:
            assert _planes_digest(out[g], cl) == want, "frame %d" % g
            checked += 1
    assert checked == N - bench.REACH


def test_bench_workload_shard_with_halo_equals_the_unsharded_clip(gpu_ctx):
    """a rank's shard (frames 37 .. 59 of a longer clip, REACH halo frames on each side) gives the frames of the whole clip"""
    gold = bench.load_golden("headline")
    first, count = 37, 23
    cfg, sp, aps, dp, cl, host, lo, F, _ = _setup(first, count, 200)
    L = mb.lib()
    apa = (C.POINTER(mb.AnalyseParams) * len(aps))(*[C.pointer(a) for a in aps])
    out = np.zeros((count, cl.frame_bytes), dtype=np.uint8)
    mb._check(L.mvb200_degrain_clip_host_async(gpu_ctx.handle, C.byref(sp), apa, len(aps), C.byref(dp), host.ctypes.data, out.ctypes.data, F, first - lo, count))
    gpu_ctx.synchronize()
    for k in range(count):
        assert _planes_digest(out[k], cl) == bench.expected_digest(gold, first + k, 200), "frame %d" % (first + k)


def test_bench_workload_through_the_plugin_default_configuration(built, monkeypatch):
    """the .vpy-shaped script through libmvtools.so with NO MVB200_* variable set (automatic look-ahead windows, asynchronous
    engine calls), frames pulled by 8 host threads, then read back one by one"""
    import minivs
    for k in list(os.environ):
        if k.startswith("MVB200_"):
            monkeypatch.delenv(k)
    gold = bench.load_golden("headline")
    cfg = bench.CONFIGS["headline"]
    total = 340  # more than one automatic window (296 frames at this configuration)
    core = minivs.Core()
    mv = core.load(mb.PLUGIN_PATH)
    out, keep = bench.build_script(mv, core, cfg, bench.make_frames(cfg, total))
    out.set_cache(total)
    before = keep[1].computed()
    out.render(0, total, 8)
    bad = [g for g in list(range(0, 70)) + list(range(280, 337)) if bench.frame_digest(cfg, out.get_frame(g)) != bench.expected_digest(gold, g, total)]
    assert not bad, "frames %s differ from the reference" % bad
    # the Super / Analyse nodes were never asked for a frame beyond what the filters' creation read
    assert keep[1].computed() == before


def test_two_rank_shards_equal_the_unsharded_run():
    """two NCCL ranks, each denoising its half of one clip with a halo: every frame equals the reference-derived digest of the
    UNSHARDED clip (needs two GPUs; gpurun --gpus 2)"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29577",
                        os.path.join(bench.ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "1", "--frames", "148", "--no-cpu-baseline"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["output_check"]["mismatches"] == 0 and line["output_check"]["frames_checked_per_rank"] >= 8


# ------------------------------------------------------------------ field shifts and static reference frames on the GPU
def _ref_or_skip():
    import minivs
    if not os.path.exists(minivs.REF_PLUGIN):
        pytest.skip("oracle/_ref/libmvtools_ref.so did not travel to this box")
    return minivs


@pytest.mark.parametrize("lookahead", ["0", None])
@pytest.mark.parametrize("tff", [None, 1])
def test_field_shifts_on_the_gpu_match_the_reference_plugin(built, monkeypatch, lookahead, tff):
    """fields=1 with an odd delta gives a non-zero fieldShift (MVAnalyse.c:136-179, MVRecalculate.c:168-174,
    MVCompensate.c:188-223, MVFlow.cpp:268-299): mv.Analyse / mv.Recalculate / mv.Compensate / mv.Flow through the PRODUCT
    plugin and the CUDA engine against the compiled reference plugin, with the parity from the _Field property and from tff,
    node by node and with the default look-ahead"""
    import pipeline
    minivs = _ref_or_skip()
    if lookahead is None:
        monkeypatch.delenv("MVB200_LOOKAHEAD", raising=False)
    else:
        monkeypatch.setenv("MVB200_LOOKAHEAD", lookahead)
    case = "blk8_ov4_pel2_p8"
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    n_frames = len(frames)
    props = [{"_Field": (n + 1) % 2} for n in range(n_frames)] if tff is None else None
    fkw = dict(fields=1) if tff is None else dict(fields=1, tff=tff)
    outs = []
    for path in (minivs.REF_PLUGIN, mb.PLUGIN_PATH):
        core = minivs.Core()
        mv = core.load(path)
        clip = core.source(frames, bits=8, props=props)
        sup = mv.Super(clip, **skw)
        vb = mv.Analyse(sup, isb=1, delta=1, **fkw, **akw)
        vf = mv.Analyse(sup, isb=0, delta=1, **fkw, **akw)
        comp = mv.Compensate(clip, sup, vb, **fkw)
        flow = mv.Flow(clip, sup, vf, time=60.0, **fkw)
        rec = mv.Recalculate(sup, vb, blksize=8, overlap=4, thsad=50, **fkw)
        res = []
        for n in range(n_frames):
            res.append((vb.get_frame(n).prop_data("MVTools_vectors"), vf.get_frame(n).prop_data("MVTools_vectors"), rec.get_frame(n).prop_data("MVTools_vectors"),
                        [pipeline.sha(p) for p in comp.get_frame(n).planes()], [pipeline.sha(p) for p in flow.get_frame(n).planes()]))
        outs.append(res)
    for n in range(n_frames):
        for k, what in enumerate(("backward vectors", "forward vectors", "recalculated vectors", "Compensate", "Flow")):
            assert outs[0][n][k] == outs[1][n][k], "frame %d: %s" % (n, what)
    assert any(outs[0][n][0] != outs[0][(n + 1) % n_frames][0] for n in range(n_frames))


def test_static_reference_frame_on_the_gpu_matches_the_reference_plugin(built, monkeypatch):
    """delta < 0 (every frame searched against ONE fixed frame, MVAnalyse.c:84-109) through the product plugin with its default
    look-ahead, against the compiled reference plugin"""
    import pipeline
    minivs = _ref_or_skip()
    monkeypatch.delenv("MVB200_LOOKAHEAD", raising=False)
    case = "blk8_ov4_pel2_p8"
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    outs = []
    for path in (minivs.REF_PLUGIN, mb.PLUGIN_PATH):
        core = minivs.Core()
        mv = core.load(path)
        clip = core.source(frames, bits=8)
        sup = mv.Super(clip, **skw)
        vec = mv.Analyse(sup, isb=0, delta=-2, **akw)
        comp = mv.Compensate(clip, sup, vec)
        rec = mv.Recalculate(sup, vec, blksize=8, overlap=4, thsad=50)
        outs.append([(vec.get_frame(n).prop_data("MVTools_vectors"), rec.get_frame(n).prop_data("MVTools_vectors"), [pipeline.sha(p) for p in comp.get_frame(n).planes()])
                     for n in range(len(frames))])
    assert outs[0] == outs[1]


@pytest.mark.parametrize("lookahead", ["3", None])
def test_recalculate_passes_inside_the_fused_chains_on_the_gpu(built, monkeypatch, lookahead):
    """Analyse -> Recalculate (-> Recalculate) -> Degrain1 / Compensate through the product plugin (fused: one engine call per
    window, mvb200_degrain_recalc_clip_host_async / mvb200_compensate_recalc_clip_host) against the reference plugin"""
    import pipeline
    minivs = _ref_or_skip()
    if lookahead:
        monkeypatch.setenv("MVB200_LOOKAHEAD", lookahead)
    else:
        monkeypatch.delenv("MVB200_LOOKAHEAD", raising=False)
    case = "blk16_ov8_pel2_p8_r3"
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    outs, counts = [], None
    for path in (minivs.REF_PLUGIN, mb.PLUGIN_PATH):
        core = minivs.Core()
        mv = core.load(path)
        clip = core.source(frames, bits=8)
        sup = mv.Super(clip, **skw)
        vb, vf = mv.Analyse(sup, isb=1, delta=1, **akw), mv.Analyse(sup, isb=0, delta=1, **akw)
        rb = mv.Recalculate(sup, vb, blksize=8, overlap=4, thsad=100)
        rb2 = mv.Recalculate(sup, rb, blksize=8, overlap=4, thsad=50, search=3, searchparam=1)
        rf = mv.Recalculate(sup, vf, blksize=8, overlap=4, thsad=100, divide=0)
        dn = mv.Degrain1(clip, sup, rb2, rf, thsad=300)
        comp = mv.Compensate(clip, sup, rb)
        nodes = (sup, vb, vf, rb, rb2, rf)
        before = [x.computed() for x in nodes]
        outs.append([([pipeline.sha(p) for p in dn.get_frame(n).planes()], [pipeline.sha(p) for p in comp.get_frame(n).planes()]) for n in range(len(frames))])
        if path == mb.PLUGIN_PATH:
            counts = (before, [x.computed() for x in nodes])
    assert outs[0] == outs[1]
    assert counts[0] == counts[1], "the fused path asked the Super / Analyse / Recalculate nodes for frames"


@pytest.mark.parametrize("lookahead", ["3", None])
@pytest.mark.parametrize("case", ["c1_640x360_p8_blk8_pel1", "blk16_ov8_pel2_p16_r3", "gray_blk4_pel1_rf0", "divide2_blk16_esa_neg_badrange", "super_nochroma_blk8_r2", "p8_420_blk8_pel4"])
def test_vector_clip_windows_from_source_frames_on_the_gpu(built, monkeypatch, case, lookahead):
    """a vector clip read directly (here also by mv.SCDetection): windows of vector frames from ONE mvb200_analyse_clip_host call over
    source frames -- pixels (the super frame), vectors and properties against the goldens"""
    import json
    import minivs
    import pipeline
    gold = json.load(open(os.path.join(bench.ROOT, "tests", "golden", "golden.json")))
    if lookahead:
        monkeypatch.setenv("MVB200_LOOKAHEAD", lookahead)
    else:
        monkeypatch.delenv("MVB200_LOOKAHEAD", raising=False)
    ck, skw, akw, radius, dkw = pipeline.CASES[case]
    frames = pipeline.make_frames(ck)
    n_frames = len(frames)
    sp, aps, dp = pipeline.params_for(case)
    core = minivs.Core()
    mv = core.load(mb.PLUGIN_PATH)
    clip = core.source(frames, bits=ck.get("bits", 8), ssw=ck.get("ssw", 1), ssh=ck.get("ssh", 1))
    sup = mv.Super(clip, **skw)
    vec = mv.Analyse(sup, isb=1, delta=1, **pipeline._fix(akw))
    scd = mv.SCDetection(clip, vec)
    vec.set_cache(0)
    for n in list(range(n_frames)) + [n_frames - 1, 0, n_frames // 2]:
        fr = vec.get_frame(n)
        assert [pipeline.sha(p) for p in fr.planes()] == gold[case]["super"][n], "frame %d: pixels" % n
        blob = np.frombuffer(fr.prop_data("MVTools_vectors"), dtype=np.uint8)
        assert pipeline.sha(pipeline.mask_blob(blob, aps[0])) == gold[case]["vectors"][n][0], "frame %d: vectors" % n
        assert (fr.prop_int("Super_pel") is not None) == (n == 0)
        scd.get_frame(n)
    assert sup.computed() <= 2
