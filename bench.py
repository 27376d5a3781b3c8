#!/usr/bin/env python
"""bench.py -- frames/sec of the mv.Super -> mv.Analyse (-> mv.Recalculate) -> mv.DegrainN / mv.Compensate hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME] [--frames F] [--impl reference]

The default workload (the one the driver runs) is BASELINE.json's metric: mv.Super(pel=2) + 6x mv.Analyse(blksize=16,
overlap=8, delta=1..3, isb=0/1) + mv.Degrain3 on 1920x1080 YUV420P8.  `--config c1..c5` selects BASELINE.json's
`configs[0..4]` instead (each prints the same kind of line).  One "step" = one pass of the whole path over a batch of F
synthetic frames per GPU.  Prints ONE JSON line:

  value     whole-job frames/sec with the clip resident in HBM (CUDA events on the engine's stream, max over ranks)
  e2e       the same metric through the DROP-IN PLUGIN: libmvtools.so loaded by a VapourSynth-API-4 host (host/minivs --
            this image has no VapourSynth), the .vpy-shaped script built on host frames, the host's threads pulling output
            frames -- exactly what `--impl reference` does with the reference plugin.  Default environment (no MVB200_*
            variable).  Source frames go host -> device and denoised frames device -> host inside the timed region.
  e2e.c_abi the clip-level C-ABI call with pinned host buffers (mvb200_degrain_clip_host_async), round 1's e2e
  roofline  the dominant kernel (the search): algorithmic bytes per launch / its CUDA-event time against the measured HBM peak
  cpu_baseline  the reference's own CPU code on this box's host cores (bounded sample; N=1 only)
  output_check  sampled output frames of the plugin leg compared with sha256 digests made from the compiled REFERENCE
                (tests/golden/bench_digest.json, tests/golden/make_golden_bench.py); every rank checks its own shard

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libmvtools_ref.so = the reference sources compiled in place)
under the same host with one worker thread per host core; it is the only place besides the cpu_baseline leg where
anything under oracle/ is executed by this file.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in ("vapoursynth-mvtools_b200", os.path.join("host", "minivs")):
    sys.path.insert(0, os.path.join(ROOT, _p))

import numpy as np  # noqa: E402

METRIC = "frames/sec Analyse+Degrain3 1080p YUV420P8"
DISTINCT = 64  # a 64-frame synthetic clip (pan + moving rectangle + noise, one scene cut at frame 32); longer batches cycle through it
REACH = 3      # frames on either side an output frame can depend on (radius 3, delta <= 3): the halo of a shard

# kind: "degrain" = Super + 2r Analyse + DegrainN; "analyse" = Super + one Analyse; "mc" = Super + Analyse + Recalculate + Compensate.
# frames = frames per step per GPU (a whole number of waves of the search kernel's chain slots where memory allows).
CONFIGS = {
    "headline": dict(w=1920, h=1080, bits=8, kind="degrain", radius=3, blk=16, ov=8, pel=2, frames=1184, metric=METRIC,
                     workload="mv.Super(pel=2) + 6x mv.Analyse(blksize=16, overlap=8, delta=1..3, isb=0/1) + mv.Degrain3, 1920x1080 YUV420P8 "
                              "(metric's Analyse+Degrain3 1080p P8 with configs[2]'s block geometry)"),
    "c1": dict(w=640, h=360, bits=8, kind="analyse", blk=8, ov=0, pel=1, search=4, frames=2368, distinct=30, metric="frames/sec Super+Analyse 640x360 YUV420P8",
               workload="configs[0]: mv.Super(pel=1) + mv.Analyse(blksize=8, pel=1, search=4), 640x360 YUV420P8, the 30-frame clip cycled"),
    "c2": dict(w=1920, h=1080, bits=8, kind="degrain", radius=1, blk=8, ov=4, pel=2, frames=3072, metric="frames/sec Analyse+Degrain1 1080p YUV420P8",
               workload="configs[1]: mv.Super(pel=2) + 2x mv.Analyse(blksize=8, overlap=4) + mv.Degrain1, 1920x1080 YUV420P8"),
    "c3": dict(w=1920, h=1080, bits=16, kind="degrain", radius=3, blk=16, ov=8, pel=2, frames=888, metric="frames/sec Analyse+Degrain3 1080p YUV420P16",
               workload="configs[2]: mv.Super(pel=2) + 6x mv.Analyse(blksize=16, overlap=8, delta=1..3) + mv.Degrain3, 1920x1080 YUV420P16"),
    "c4": dict(w=3840, h=2160, bits=8, kind="mc", blk=8, ov=0, pel=2, dct=5, frames=296, distinct=32, metric="frames/sec Analyse(SATD)+Recalculate+Compensate 2160p YUV420P8",
               workload="configs[3]: mv.Super(pel=2) + mv.Analyse(isb=1, dct=5 = SATD) + mv.Recalculate(dct=5) + mv.Compensate, 3840x2160 YUV420P8 (block defaults: 8x8, no overlap)"),
    "c5": dict(w=3840, h=2160, bits=16, kind="degrain", radius=3, blk=16, ov=8, pel=2, frames=512, fixed_total=512, metric="frames/sec Analyse+Degrain3 2160p YUV420P16, 512-frame batch",
               workload="configs[4]: mv.Super(pel=2) + 6x mv.Analyse(blksize=16, overlap=8, delta=1..3) + mv.Degrain3, 3840x2160 YUV420P16, a FIXED 512-frame batch sharded across the GPUs"),
}

_BASE_CLIPS = {}


def make_frames(cfg, n, first=0):
    """Frames [first, first + n) of the config's cyclic synthetic clip (list of [Y, U, V]); repeated frames are the same objects."""
    import clipgen
    distinct = cfg.get("distinct", DISTINCT)
    key = (cfg["w"], cfg["h"], cfg["bits"], distinct)
    if key not in _BASE_CLIPS:
        _BASE_CLIPS[key] = clipgen.make_clip(cfg["w"], cfg["h"], distinct, bits=cfg["bits"], seed=0)
    base = _BASE_CLIPS[key]
    return [base[(first + i) % distinct] for i in range(n)]


def build_script(mv, core, cfg, frames):
    """The .vpy-shaped script of a config on host frames, through whichever plugin `mv` is.  Returns (output node, keep-alive list)."""
    clip = core.source(frames, bits=cfg["bits"])
    sup = mv.Super(clip, pel=cfg["pel"])
    akw = dict(blksize=cfg["blk"], overlap=cfg["ov"])
    for k in ("search", "dct"):
        if k in cfg:
            akw[k] = cfg[k]
    if cfg["kind"] == "degrain":
        r = cfg["radius"]
        vecs = [mv.Analyse(sup, isb=isb, delta=d, **akw) for d in range(1, r + 1) for isb in (1, 0)]
        out = mv.invoke("Degrain%d" % r, clip, sup, *vecs)
        return out, [clip, sup] + vecs
    if cfg["kind"] == "analyse":
        vec = mv.Analyse(sup, **akw)
        return vec, [clip, sup]
    vec = mv.Analyse(sup, isb=1, **akw)
    rec = mv.Recalculate(sup, vec, blksize=cfg["blk"], overlap=cfg["ov"], dct=cfg.get("dct", 0))
    out = mv.Compensate(clip, sup, rec)
    return out, [clip, sup, vec, rec]


def frame_digest(cfg, fr):
    """sha256 of an output frame: its planes, or -- for a vector clip -- its MVTools_vectors blob."""
    h = hashlib.sha256()
    if cfg["kind"] == "analyse":
        h.update(fr.prop_data("MVTools_vectors"))
    else:
        for p in fr.planes():
            h.update(np.ascontiguousarray(p).tobytes())
    return h.hexdigest()


def expected_digest(gold, g, total):
    """Digest of global frame g of a cyclic clip of `total` frames, from the head / cycle digests of the golden file
    (frames closer than REACH to the clip's end depend on the clip length: not pinned -> None)."""
    if gold is None or g >= total - REACH:
        return None
    if g < REACH:
        return gold["head"][g]
    return gold["cycle"][g % len(gold["cycle"])]


def load_golden(name):
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "bench_digest.json"))).get(name)
    except (OSError, ValueError):
        return None


# ----------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0]))
            mx = max(mx, int(f[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * mx] or sm  # (idle samples between the legs of the run are not "under load")
        return {"sm_mhz": int(np.median(busy)) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------- reference (CPU) arm
def reference_fps(cfg, sample_frames, warm_frames, threads):
    """Frames/sec of the reference's own CPU path on `sample_frames` frames of the workload, after `warm_frames` untimed ones."""
    import minivs
    if not os.path.exists(minivs.REF_PLUGIN):
        return None, "oracle/_ref/libmvtools_ref.so is not built"
    total = warm_frames + sample_frames
    core = minivs.Core()
    mv = core.load(minivs.REF_PLUGIN)
    out, keep = build_script(mv, core, cfg, make_frames(cfg, total))
    keep[1].set_cache(64 + 2 * threads)  # the super clip: every vector clip and the consumer read it
    for v in keep[2:]:
        v.set_cache(8 + 2 * threads)
    out.set_cache(0)
    if warm_frames:
        out.render(0, warm_frames, threads)
    t = out.render(warm_frames, total, threads)
    return sample_frames / t, None


def reference_sample(cfg):
    """Frames per reference step: about 10-30 s of CPU work on a 16-core box, a whole cycle of the clip where that fits."""
    px = cfg["w"] * cfg["h"] * (2 if cfg["bits"] > 8 else 1)
    return cfg.get("distinct", DISTINCT) if px <= 1920 * 1080 else 32


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = reference_sample(cfg)
    times = []
    for _ in range(args.warmup):
        reference_fps(cfg, max(4, sample // 4), 0, cores)
    for _ in range(args.steps):
        fps, err = reference_fps(cfg, sample, 4, cores)
        if fps is None:
            print(json.dumps({"impl": "reference", "unavailable": err}))
            return
        times.append(sample / fps)
    fps = sample * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "strong" if cfg.get("fixed_total") else "weak", "vs_baseline": None, "dtype": "u8" if cfg["bits"] <= 8 else "u16", "data": "synthetic",
        "config": {"workload": cfg["workload"], "name": args.config, "frames_per_step": sample, "host_threads": cores},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                         "sample": "%d frames per step after 4 warm-up frames, all host cores, the reference plugin built from its own sources and driven by the same "
                                   "API-4 host as the product plugin (x264 NASM kernels replaced by the reference's C fallbacks: u8 4x4 / 4x8 SAD and u8 SATD)" % sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm
def fill_host_clip(hv, frames, cl):
    for n, fr in enumerate(frames):
        for p, a in enumerate(fr):
            rows = hv[n, cl.offset[p]: cl.offset[p] + cl.pitch[p] * cl.plane_h[p]].reshape(cl.plane_h[p], cl.pitch[p])
            rows[:, : a.shape[1] * a.itemsize] = a.view(np.uint8).reshape(a.shape[0], -1)


def plugin_window(mb, aps, sp, kind, device):
    """(frames per window, engine contexts) of the plugin's fused paths in its default configuration on one device: the engine's own answer
    (mvb200_window_plan_query, which plugin.cpp fusedWindowFrames calls with the same arguments) -- used to size the resident
    leg's windows and to ALIGN the timed region of the plugin leg, see plugin_leg."""
    L = mb.lib()
    cl, sl = mb.clip_layout(sp), mb.super_layout(sp)
    out_bytes = sl.frame_bytes + (aps[0].vectors_size + 255) // 256 * 256 if kind == "analyse" else cl.frame_bytes
    per = int(os.environ.get("MVB200_CONTEXTS_PER_DEVICE", "0") or 0)  # (unset: the plan chooses)
    wp = mb.WindowPlan()
    mb._check(L.mvb200_window_plan_query(C.byref(aps[0]), C.byref(sp), len(aps), -1, cl.frame_bytes, out_bytes, 1, per, 0, L.mvb200_device_memory(device), C.byref(wp)))
    return int(wp.frames), int(wp.contexts_per_device)


def plugin_leg(cfg, args, mb, plan, world, rank, local, barrier, total_frames, gold, window, contexts=2):
    """e2e through libmvtools.so under the API-4 host, default environment.
    Weak-scaling configs: STEADY STATE -- one long clip; W warm-up steps, then K timed steps.  The plugin works in aligned
    windows and keeps the next window in flight, so a timed region that started or ended inside a window would inherit (or
    give away) part of a window's work: the step is rounded up to whole windows, the region starts on a window boundary
    and the clip continues for a window past it.
    A fixed batch (c5): COLD START -- the engine is warmed on another instance of the script, then a fresh instance renders
    exactly the batch, first request to last frame (pipeline fill and drain included)."""
    import minivs
    os.environ["MVB200_DEVICE"] = str(local)
    F_out, OFF, lo = plan["count"], plan["offset"], plan["lo"]
    threads = max(4, (os.cpu_count() or 4) // world)

    def script(n_local, first):
        core = minivs.Core()
        mv = core.load(mb.PLUGIN_PATH)
        out, keep = build_script(mv, core, cfg, make_frames(cfg, n_local, first=first))
        out.set_cache(0)  # (every config's consumer is computed in windows from the source frames: nothing upstream is read through a cache)
        return core, out, keep

    if cfg.get("fixed_total"):
        tail = plan["local_frames"] - OFF - F_out
        nwarm = min(F_out, (contexts + 3) * window)  # (every context's scratch and the pinned windows in circulation: both are per process)
        warm = script(OFF + nwarm + tail, lo)
        warm[1].render(OFF, OFF + nwarm, threads)
        del warm
        core, out, keep = script(plan["local_frames"], lo)
        barrier()
        t = out.render(OFF, OFF + F_out, threads)
        barrier()
        frames_timed, first_timed, protocol = F_out, OFF, "cold start: a fresh script instance renders the whole batch once (engine warmed on another instance)"
    else:
        step = -(-F_out // window) * window
        base = -(-OFF // window) * window  # the first rendered frame: a window boundary at or after the halo
        # warm-up: the requested steps, and at least the windows it takes for the plugin's pinned buffers to be in circulation (one
        # per engine context, one being served, one recycled, one more); then window by window until two consecutive windows take
        # the same time within 25 % (page-locking the windows' gigabytes is a start-up cost of seconds: a small configuration
        # renders a dozen windows before the last buffer exists), at most 24 more
        warm_min = max(args.warmup_e2e * step, (contexts + 3) * window) if args.warmup_e2e else 0
        warm_min = -(-warm_min // window) * window
        warm_max = warm_min + (24 * window if args.warmup_e2e else 0)
        # timed: the K steps, and at least four times the windows the plugin keeps in flight (they straddle both ends of the
        # region: a region of four windows with four in flight measured 596 - 803 frames/s on the same box, run to run)
        timed = max(args.steps * step, 4 * contexts * window)
        n_local = base + warm_max + timed + window + REACH
        core, out, keep = script(n_local, lo)
        pos, prev = base, None
        if warm_min > window:
            out.render(base, base + warm_min - window, threads)
            pos = base + warm_min - window
        while pos < base + warm_max:
            tw = out.render(pos, pos + window, threads)
            pos += window
            if pos - base >= warm_min and prev is not None and abs(tw - prev) <= 0.25 * max(tw, prev):
                break
            prev = tw
        warm = pos - base
        barrier()
        t = out.render(pos, pos + timed, threads)
        barrier()
        frames_timed, first_timed = timed, pos
        protocol = "steady state: %d warm-up frames (%d windows, until two consecutive windows took the same time), then %d timed frames (%d steps of %d frames, at least 4 x %d contexts x windows of %d) on one long clip" % (
            warm, warm // window, timed, args.steps, step, contexts, window)
    # ---- output check (untimed): sampled frames against reference-derived digests
    checked, bad = 0, []
    if gold is not None:
        cyc = cfg.get("distinct", DISTINCT)
        picks = sorted(set(first_timed + k for k in (0, 1, 2, cyc // 2 - 2, cyc // 2 - 1, cyc // 2, cyc // 2 + 1, cyc - 2, cyc - 1, cyc, cyc + 1) if k < F_out))
        for n in picks:
            want = expected_digest(gold, lo + n, 1 << 30)
            if want is None:
                continue
            checked += 1
            if frame_digest(cfg, out.get_frame(n)) != want:
                bad.append(lo + n)
    return dict(seconds=t, frames=frames_timed, threads=threads, checked=checked, bad=bad, protocol=protocol)


def run_gpu_arm(args, cfg):
    import torch
    import torch.distributed as dist
    import mvtools_b200 as mb
    import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = mb.lib()
    W_, H_, BITS, kind = cfg["w"], cfg["h"], cfg["bits"], cfg["kind"]
    # Frames are independent units: a global clip is split into contiguous ranges; a rank holds its range plus a
    # REACH-frame halo of source frames on each side (their super frames are recomputed locally, no data-path collective).
    # Weak scaling: world x F frames.  A config with fixed_total (c5) is a STRONG-scaling run: the same batch whatever N.
    total = cfg["fixed_total"] if cfg.get("fixed_total") else world * args.frames
    plan = sharding.local_plan(total, world, rank, reach=REACH)
    F_out, F, OFF = plan["count"], plan["local_frames"], plan["offset"]
    sp = mb.super_params(W_, H_, BITS, 1, 1, pel=cfg["pel"])
    akw = dict(blksize=cfg["blk"], overlap=cfg["ov"])
    for k in ("search", "dct"):
        if k in cfg:
            akw[k] = cfg[k]
    if kind == "degrain":
        aps = [mb.analyse_params(sp, F, isb=isb, delta=d, **akw) for d in range(1, cfg["radius"] + 1) for isb in (1, 0)]
        dp = mb.degrain_params(sp, [a.ad for a in aps], W_, H_, cfg["radius"])
    elif kind == "analyse":
        aps = [mb.analyse_params(sp, F, **akw)]
    else:
        aps = [mb.analyse_params(sp, F, isb=1, **akw)]
        rp = mb.recalculate_params(sp, aps[0].ad, F, blksize=cfg["blk"], overlap=cfg["ov"], dct=cfg.get("dct", 0))
        cp = mb.compensate_params(sp, rp.ad, W_, H_, F)
    plan_q = mb.SearchPlan()
    mb._check(L.mvb200_search_plan_query(C.byref(aps[0]), C.byref(sp), 0, C.byref(plan_q)))
    cl, sl = mb.clip_layout(sp), mb.super_layout(sp)
    ctx = mb.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    frames = make_frames(cfg, F, first=plan["lo"])
    host = torch.zeros((F, cl.frame_bytes), dtype=torch.uint8).pin_memory()
    fill_host_clip(host.numpy(), frames, cl)
    vs = (max(a.vectors_size for a in aps) + 255) // 256 * 256
    dsrc = host.to("cuda")  # (torch's own stream: the pinned block must not be tied to the engine's stream, which dies with ctx)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        dsup = torch.empty((F, sl.frame_bytes), dtype=torch.uint8, device="cuda")
        dvec = [torch.empty((F_out, vs), dtype=torch.uint8, device="cuda") for _ in aps]
        ddst = torch.empty((F_out, cl.frame_bytes), dtype=torch.uint8, device="cuda") if kind != "analyse" else None
        if kind == "mc":
            rvs = (rp.vectors_size + 255) // 256 * 256
            drec = torch.empty((F_out, rvs), dtype=torch.uint8, device="cuda")
    apa = (C.POINTER(mb.AnalyseParams) * len(aps))(*[C.pointer(a) for a in aps])
    outs = (C.c_void_p * len(aps))(*[t.data_ptr() for t in dvec])

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # the stages of one step, in stream order; vectors and output only for this rank's own frames (the halo frames are references)
    stages = [("super", lambda: L.mvb200_super_batch(ctx.handle, C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), F)),
              ("analyse_x%d" % len(aps), lambda: L.mvb200_analyse_multi_batch(ctx.handle, apa, len(aps), C.byref(sp), dsup.data_ptr(), OFF, F_out, outs, vs))]
    if kind == "degrain":
        stages.append(("degrain%d" % cfg["radius"], lambda: L.mvb200_degrain_batch(ctx.handle, C.byref(dp), C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), F, OFF, F_out, outs, vs, ddst.data_ptr())))
    elif kind == "mc":
        stages.append(("recalculate", lambda: L.mvb200_recalculate_batch(ctx.handle, C.byref(rp), C.byref(sp), dsup.data_ptr(), OFF, F_out, dvec[0].data_ptr(), vs, drec.data_ptr(), rvs)))
        stages.append(("compensate", lambda: L.mvb200_compensate_batch(ctx.handle, C.byref(cp), C.byref(sp), dsup.data_ptr(), OFF, F_out, drec.data_ptr(), rvs, ddst.data_ptr())))

    def pipeline(timers=None):
        for i, (_, fn) in enumerate(stages):
            if timers is not None:
                timers[i].record(stream)
            mb._check(fn())
        if timers is not None:
            timers[len(stages)].record(stream)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing, pass 1: ONE context, stages serialised on its stream -> per-stage times (the roofline's kernel time)
    aux_steps = max(1, min(2, args.steps))
    pipeline()
    barrier()
    stage_events = [[ev() for _ in range(len(stages) + 1)] for _ in range(aux_steps)]
    for k in range(aux_steps):
        pipeline(stage_events[k])
    barrier()
    stage_ms = {name: sum(e[i].elapsed_time(e[i + 1]) for e in stage_events) / aux_steps for i, (name, _) in enumerate(stages)}
    search_ms = stage_ms[stages[1][0]]
    value_one_context = (total if cfg.get("fixed_total") else world * F_out) * aux_steps / (max_over_ranks(stage_events[0][0].elapsed_time(stage_events[-1][-1])) / 1e3)
    del dsup, dvec, ddst, outs
    if kind == "mc":
        del drec
    torch.cuda.empty_cache()

    # ---- pass 2 (the `value`): the product's execution mode -- the step's frames flow through C engine contexts on the device as
    # WINDOWS (the plugin's own window size: what fills the search kernel's chain slots once), window j on context j mod C, each
    # context running Super -> searches -> DegrainN of its windows back to back on its own stream: one window's Super / DegrainN
    # kernels run beside the other's search, and the next window's search chains take the slots the moment the slow chains of the
    # previous one leave them.  W warm-up steps, K timed steps, all
    # enqueued without a host sync in between; time = first start -> last end over all streams (CUDA events), max over ranks.
    PW, plan_contexts = plugin_window(mb, aps, sp, kind, local)
    NC = max(1, args.contexts if args.contexts else plan_contexts)
    S = max(8, min(PW, F_out))
    window_frames = S
    bounds = list(range(0, F_out, S)) + [F_out]
    ctxs = [ctx] + [mb.Context(local) for _ in range(NC - 1)]
    streams = [torch.cuda.ExternalStream(c.stream(), device=torch.device("cuda", local)) for c in ctxs]
    bufs = []
    for c_i, st_i in zip(ctxs, streams):
        with torch.cuda.stream(st_i):
            bufs.append(dict(sup=torch.empty((S + 2 * REACH, sl.frame_bytes), dtype=torch.uint8, device="cuda"),
                             vec=[torch.empty((S, vs), dtype=torch.uint8, device="cuda") for _ in aps],
                             rec=torch.empty((S, rvs), dtype=torch.uint8, device="cuda") if kind == "mc" else None))
    ddst_all = torch.empty((F_out, cl.frame_bytes), dtype=torch.uint8, device="cuda") if kind != "analyse" else None
    torch.cuda.synchronize()
    params_for = {}

    def window_params(Fi):
        if Fi not in params_for:
            if kind == "degrain":
                aps_i = [mb.analyse_params(sp, Fi, isb=isb, delta=d, **akw) for d in range(1, cfg["radius"] + 1) for isb in (1, 0)]
                extra = None
            elif kind == "analyse":
                aps_i, extra = [mb.analyse_params(sp, Fi, **akw)], None
            else:
                aps_i = [mb.analyse_params(sp, Fi, isb=1, **akw)]
                rp_i = mb.recalculate_params(sp, aps_i[0].ad, Fi, blksize=cfg["blk"], overlap=cfg["ov"], dct=cfg.get("dct", 0))
                extra = (rp_i, mb.compensate_params(sp, rp_i.ad, W_, H_, Fi))
            params_for[Fi] = (aps_i, (C.POINTER(mb.AnalyseParams) * len(aps_i))(*[C.pointer(a) for a in aps_i]), extra)
        return params_for[Fi]

    windows = []
    for j in range(len(bounds) - 1):
        a0, a1 = bounds[j], bounds[j + 1]
        lo_i, hi_i = max(0, OFF + a0 - REACH), min(F, OFF + a1 + REACH)
        windows.append(dict(slot=j % NC, Fi=hi_i - lo_i, off=OFF + a0 - lo_i, cnt=a1 - a0, src=dsrc.data_ptr() + lo_i * cl.frame_bytes, dst=a0))

    def run_window(wd, slot):
        c_i, B = ctxs[slot], bufs[slot]
        aps_i, apa_i, extra = window_params(wd["Fi"])
        outs_i = (C.c_void_p * len(aps_i))(*[t.data_ptr() for t in B["vec"]])
        mb._check(L.mvb200_super_batch(c_i.handle, C.byref(sp), wd["src"], B["sup"].data_ptr(), wd["Fi"]))
        mb._check(L.mvb200_analyse_multi_batch(c_i.handle, apa_i, len(aps_i), C.byref(sp), B["sup"].data_ptr(), wd["off"], wd["cnt"], outs_i, vs))
        if kind == "degrain":
            mb._check(L.mvb200_degrain_batch(c_i.handle, C.byref(dp), C.byref(sp), wd["src"], B["sup"].data_ptr(), wd["Fi"], wd["off"], wd["cnt"], outs_i, vs,
                                             ddst_all.data_ptr() + wd["dst"] * cl.frame_bytes))
        elif kind == "mc":
            mb._check(L.mvb200_recalculate_batch(c_i.handle, C.byref(extra[0]), C.byref(sp), B["sup"].data_ptr(), wd["off"], wd["cnt"], B["vec"][0].data_ptr(), vs, B["rec"].data_ptr(), rvs))
            mb._check(L.mvb200_compensate_batch(c_i.handle, C.byref(extra[1]), C.byref(sp), B["sup"].data_ptr(), wd["off"], wd["cnt"], B["rec"].data_ptr(), rvs,
                                                ddst_all.data_ptr() + wd["dst"] * cl.frame_bytes))

    parts = [dict(ctx=c_i, stream=st_i) for c_i, st_i in zip(ctxs, streams)]

    turn = [0]  # windows go round the contexts ACROSS steps (a step of two windows would keep to two contexts otherwise)

    def run_step():
        for wd in windows:
            run_window(wd, turn[0] % NC)
            turn[0] += 1

    def barrier_all():
        for P in parts:
            P["ctx"].synchronize()
        barrier()

    for _ in range(max(args.warmup, -(-NC // len(windows)))):  # (every context runs at least one window before the timed region)
        run_step()
    barrier_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = sum(P["ctx"].launch_count() for P in parts)
    e_start = ev()
    e_ends = [ev() for _ in parts]
    barrier_all()
    e_start.record(parts[0]["stream"])
    for k in range(args.steps):
        run_step()
    for P, e in zip(parts, e_ends):
        e.record(P["stream"])
    barrier_all()
    launches = sum(P["ctx"].launch_count() for P in parts) - launches0
    region_ms = max_over_ranks(max(e_start.elapsed_time(e) for e in e_ends))
    value = (total if cfg.get("fixed_total") else world * F_out) * args.steps / (region_ms / 1e3)
    for P in parts[1:]:
        P["ctx"].close()
    del parts, dsrc, bufs, ddst_all, windows
    torch.cuda.empty_cache()

    # ---- the clip-level C ABI with pinned host buffers (round 1's e2e; DegrainN pipelines only).  (The resident legs' device
    # buffers are gone: the engine allocates its own clips, and a 4K batch does not fit twice.)
    c_abi = None
    if kind == "degrain":
        host_outs = [torch.zeros((F_out, cl.frame_bytes), dtype=torch.uint8).pin_memory() for _ in range(2)]
        calls = [0]

        def e2e_step():
            o = host_outs[calls[0] & 1]
            calls[0] += 1
            mb._check(L.mvb200_degrain_clip_host_async(ctx.handle, C.byref(sp), apa, len(aps), C.byref(dp), host.data_ptr(), o.data_ptr(), F, OFF, F_out))

        for _ in range(2):
            e2e_step()
        barrier()
        e0, e1 = ev(), ev()
        e0.record(stream)
        for _ in range(args.steps):
            e2e_step()
        ctx.synchronize()  # compute stream and both copy streams of the engine
        e1.record(stream)
        barrier()
        c_abi_ms = max_over_ranks(e0.elapsed_time(e1))
        c_abi = (total if cfg.get("fixed_total") else world * F_out) * args.steps / (c_abi_ms / 1e3)
        del host_outs
    clocks_dev = sampler.stop() if rank == 0 else None

    # ---- e2e through the drop-in plugin (the headline); the plugin has its own context(s): this one's memory goes first
    del host
    ctx.close()
    torch.cuda.empty_cache()
    gold = load_golden(args.config)
    cl_frame_bytes, sl_frame_bytes = int(cl.frame_bytes), int(sl.frame_bytes)
    sampler2 = ClockSampler(local)
    if rank == 0:
        sampler2.start()
    def host_barrier():  # (no device synchronisation here: the plugin's windows in flight must stay in flight across the boundary
        if world > 1:    #  between the warm-up and the timed steps -- a drained pipeline would hand the timed region finished windows)
            dist.barrier()

    pl = plugin_leg(cfg, args, mb, plan, world, rank, local, host_barrier, total, gold, PW, plan_contexts)
    clocks_e2e = sampler2.stop() if rank == 0 else None
    e2e_s = max_over_ranks(pl["seconds"])
    e2e_frames = total if cfg.get("fixed_total") else world * pl["frames"]
    e2e_value = e2e_frames / e2e_s
    nbad = max_over_ranks(float(len(pl["bad"])))
    nchecked = -max_over_ranks(-float(pl["checked"]))  # the minimum over ranks
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if pl["bad"]:
            raise SystemExit("bench.py: rank %d: output frames %s differ from the reference-derived digests" % (rank, pl["bad"]))
        return

    # ---- roofline of the dominant kernel (the search): algorithmic bytes per launch
    bps = 2 if BITS > 8 else 1
    A = 1.5 * (W_ + 32) * (H_ + 32) * bps
    S = (cfg["pel"] ** 2 + 1.0 / 3.0) * A      # reference side: the whole super frame
    S1 = (1 + 1.0 / 3.0) * A                   # source side: plane 0 of every level
    alg_bytes = sum(F_out * (S1 + S + a.vectors_size) for a in aps)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (search_ms / 1e3) / 1e9
    # DRAM traffic and pipe utilisation of the same kernel from the committed ncu --set full capture (per chain, scaled to this launch)
    traffic, int_pipe, prof_name = None, None, cfg.get("ncu_profile", "r2_ncu_search_headline.json")
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", prof_name)))
        tobytes = lambda m: m["value"] * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[m["unit"]]
        if args.config == prof.get("_config", "headline"):
            traffic = (tobytes(prof["dram__bytes_read.sum"]) + tobytes(prof["dram__bytes_write.sum"])) / prof["_chains"] * len(aps) * F_out
        int_pipe = {"alu_pipe_active_pct_of_peak": prof["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]["value"],
                    "issue_active_pct_of_peak": prof["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"],
                    "source": "profiles/" + prof_name}
    except (OSError, KeyError, ValueError):
        pass
    line = {
        "metric": cfg["metric"], "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": region_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if cfg.get("fixed_total") else "weak", "vs_baseline": None, "dtype": "u8" if BITS <= 8 else "u16", "data": "synthetic",
        "config": {"workload": cfg["workload"], "name": args.config, "frames_per_step_per_gpu": F_out, "halo_frames_per_gpu": F - F_out, "distinct_frames": cfg.get("distinct", DISTINCT),
                   "search_chains_per_step_per_gpu": len(aps) * F_out,
                   "search_kernel": {0: "k_search (generic)", 1: "k_search_fast (register-blocked)", 2: "k_search_tma (TMA-staged windows)"}[plan_q.kernel],
                   "search_chain_slots_per_gpu": plan_q.chains_per_wave,
                   "l2": "inputs larger than L2 (%d MB of super frames + vectors per step vs 126 MB L2), no flush" % int((F * sl.frame_bytes + len(aps) * F_out * vs) / 1e6),
                   "contexts_per_gpu": NC, "window_frames": window_frames, "value_one_context": value_one_context,
                   "stage_ms": stage_ms, "stage_ms_note": "one context, the whole step in one launch per stage, serialised on one stream (pass 1); `value` streams windows of %d frames through %d contexts" % (window_frames, NC)},
        "e2e": {"value": e2e_value, "unit": "frames/s", "through": "libmvtools.so (VapourSynthPluginInit2) under host/minivs, default environment, %d host threads per rank" % pl["threads"],
                "timer": "wall clock around the host's render of K steps (the plugin owns its CUDA context), max over ranks",
                "protocol": pl["protocol"],
                "h2d_bytes_per_step": int(pl["frames"] // max(args.steps, 1) * cl_frame_bytes) if not cfg.get("fixed_total") else int(F_out * cl_frame_bytes),
                "d2h_bytes_per_step": (int(pl["frames"] // max(args.steps, 1) * cl_frame_bytes) if not cfg.get("fixed_total") else int(F_out * cl_frame_bytes)) if kind != "analyse" else int(F_out * (aps[0].vectors_size + sl_frame_bytes)),
                "seconds": e2e_s, "frames": e2e_frames, "warmup_steps": args.warmup_e2e, "c_abi": c_abi,
                "c_abi_note": "mvb200_degrain_clip_host_async with pinned host buffers, CUDA events (round 1's e2e)" if c_abi else None},
        "output_check": {"frames_checked_per_rank": int(nchecked), "mismatches": int(nbad),
                         "against": "sha256 of the compiled reference's output (tests/golden/bench_digest.json)" if gold is not None else "no golden digest for this config"},
        "gpu_launches": int(launches),
        "clocks": clocks_dev, "clocks_e2e": clocks_e2e,
        "roofline": {"kernel": "%s (one warp per search chain; %d analyse instances x F frames)" % ({0: "k_search", 1: "k_search_fast", 2: "k_search_tma"}[plan_q.kernel], len(aps)),
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": ("dram__bytes_read+write of profiles/%s, per chain x chains of this launch" % prof_name) if traffic else None,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                     "algorithmic_bytes_per_launch": alg_bytes, "int_pipe": int_pipe,
                     "note": "the search is a serial dependency chain per (frame, direction, delta): issue / integer-pipe bound (int_pipe), far below the HBM roofline by construction"},
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = reference_sample(cfg)
        reference_fps(cfg, max(4, sample // 4), 0, cores)  # the same protocol as `--impl reference`: one warm-up pass, then the sample
        fps, err = reference_fps(cfg, sample, 4, cores)
        if fps is not None:
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                                    "sample": "%d frames of the same workload after a warm-up pass and 4 warm-up frames, one worker thread per host core, reference built from "
                                              "its own sources (x264 NASM kernels replaced by the reference's C fallbacks); same protocol as --impl reference" % sample}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": "unavailable: " + err}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    torch.cuda.synchronize()
    if nbad:
        raise SystemExit("bench.py: %d sampled output frame(s) differ from the reference-derived digests" % int(nbad))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--warmup-e2e", dest="warmup_e2e", type=int, default=None, help="warm-up steps of the plugin leg (default: min(warmup, 3))")
    ap.add_argument("--config", default="headline", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=None, help="frames per step per GPU (default: the config's)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--contexts", type=int, default=0, help="engine contexts per GPU of the device-resident leg (0: what the engine's window plan says, the plugin's own choice)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.frames is None:
        args.frames = cfg["frames"]
    if args.warmup_e2e is None:
        args.warmup_e2e = min(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        run_gpu_arm(args, cfg)


if __name__ == "__main__":
    main()
