#!/usr/bin/env python
"""bench.py -- frames/sec of the mv.Super -> 6x mv.Analyse -> mv.Degrain3 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl reference]

One "step" = one pass of the whole hot path over a batch of F synthetic 1920x1080
YUV420P8 frames (per GPU): Super(pel=2) for every frame, the six Analyse searches
(blksize=16, overlap=8, delta=1..3, backward+forward) and Degrain3.  Prints ONE JSON
line (contract in the task statement): `value` = whole-job frames/sec with the clip
resident in HBM (CUDA events on the launching stream, max over ranks), `e2e` = the same
through host buffers (pinned host source -> H2D -> pipeline -> D2H of the denoised
frames inside the timed region), `roofline` for the dominant kernel (the search),
`cpu_baseline` = the reference's own CPU code on the host cores (bounded sample).

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libmvtools_ref.so, the
reference sources compiled in place, driven by the test-only mini VapourSynth host with
one worker thread per host core) on the same workload; it is the only place besides the
cpu_baseline leg where anything under oracle/ is executed by this file.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in ("vapoursynth-mvtools_b200",):
    sys.path.insert(0, os.path.join(ROOT, _p))

import numpy as np  # noqa: E402

WIDTH, HEIGHT, BITS = 1920, 1080, 8
RADIUS, BLK, OVERLAP, PEL = 3, 16, 8, 2
DISTINCT = 64  # a 64-frame synthetic clip (pan + moving rectangle + noise, one scene cut at frame 32); longer batches cycle through it
WORKLOAD = ("mv.Super(pel=2) + 6x mv.Analyse(blksize=16, overlap=8, delta=1..3, isb=0/1) + mv.Degrain3, "
            "1920x1080 YUV420P8 (metric's Analyse+Degrain3 1080p P8 with configs[2]'s block geometry)")


NCU_PROFILE = "r1_ncu_search_v7_1776chains.json"  # ncu --set full capture of the search kernel the traffic / pipe numbers come from


_BASE_CLIP = None


def make_frames(n):
    global _BASE_CLIP
    if _BASE_CLIP is None:
        import clipgen
        _BASE_CLIP = clipgen.make_clip(WIDTH, HEIGHT, DISTINCT, bits=BITS, seed=0)
    return [_BASE_CLIP[i % DISTINCT] for i in range(n)]


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
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------- reference (CPU) arm
def reference_fps(sample_frames, warm_frames, threads):
    """Frames/sec of the reference's own CPU path on `sample_frames` frames of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "minivs"))
    import minivs
    if not os.path.exists(minivs.REF_PLUGIN):
        return None, "oracle/_ref/libmvtools_ref.so is not built"
    total = warm_frames + sample_frames
    frames = make_frames(total)
    core = minivs.Core()
    mv = core.load(minivs.REF_PLUGIN)
    clip = core.source(frames, bits=BITS)
    sup = mv.Super(clip, pel=PEL)
    sup.set_cache(64 + 2 * threads)
    vecs = [mv.Analyse(sup, isb=isb, delta=d, blksize=BLK, overlap=OVERLAP) for d in range(1, RADIUS + 1) for isb in (1, 0)]
    for v in vecs:
        v.set_cache(8 + 2 * threads)
    out = mv.invoke("Degrain%d" % RADIUS, clip, sup, *vecs)
    out.set_cache(0)
    if warm_frames:
        out.render(0, warm_frames, threads)
    t = out.render(warm_frames, total, threads)
    return sample_frames / t, None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = DISTINCT  # one whole cycle of the synthetic clip (both of its discontinuities: the scene cut and the wrap-around)
    times = []
    for _ in range(args.warmup):
        reference_fps(max(4, sample // 4), 0, cores)
    for _ in range(args.steps):
        fps, err = reference_fps(sample, 4, cores)
        if fps is None:
            print(json.dumps({"impl": "reference", "unavailable": err}))
            return
        times.append(sample / fps)
    fps = sample * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": "frames/sec Analyse+Degrain3 1080p YUV420P8", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": sample, "host_threads": cores},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                         "sample": "%d frames per step after 4 warm-up frames, all host cores, reference built from its own sources "
                                   "(x264 NASM kernels replaced by the reference's C fallbacks: only u8 4x4/4x8 SAD affected here)" % sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import mvtools_b200 as mb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import sharding
    L = mb.lib()
    # weak scaling: a global clip of world x `frames` frames, split into contiguous ranges; a rank holds its
    # range plus a radius*delta halo of source frames on each side (recomputed locally, no data-path collective)
    plan = sharding.local_plan(world * args.frames, world, rank, reach=RADIUS * 1)
    F_out = plan["count"]
    F = plan["local_frames"]
    OFF = plan["offset"]
    sp = mb.super_params(WIDTH, HEIGHT, BITS, 1, 1, pel=PEL)
    aps = [mb.analyse_params(sp, F, isb=isb, delta=d, blksize=BLK, overlap=OVERLAP) for d in range(1, RADIUS + 1) for isb in (1, 0)]
    dp = mb.degrain_params(sp, [a.ad for a in aps], WIDTH, HEIGHT, RADIUS)
    cl, sl = mb.clip_layout(sp), mb.super_layout(sp)
    ctx = mb.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    frames = make_frames(world * args.frames)[plan["lo"]:plan["hi"]]
    host = torch.zeros((F, cl.frame_bytes), dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    for n, fr in enumerate(frames):
        for p, a in enumerate(fr):
            rows = hv[n, cl.offset[p]: cl.offset[p] + cl.pitch[p] * cl.plane_h[p]].reshape(cl.plane_h[p], cl.pitch[p])
            rows[:, : a.shape[1] * a.itemsize] = a.view(np.uint8).reshape(a.shape[0], -1)
    host_out = torch.zeros((F, cl.frame_bytes), dtype=torch.uint8).pin_memory()
    vs = (aps[0].vectors_size + 255) // 256 * 256
    dsrc = host.to("cuda")  # (torch's own stream: the pinned block must not be tied to the engine's stream, which dies with ctx)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        dsup = torch.empty((F, sl.frame_bytes), dtype=torch.uint8, device="cuda")
        dvec = [torch.empty((F, vs), dtype=torch.uint8, device="cuda") for _ in aps]
        ddst = torch.empty((F, cl.frame_bytes), dtype=torch.uint8, device="cuda")
    apa = (C.POINTER(mb.AnalyseParams) * len(aps))(*[C.pointer(a) for a in aps])
    outs = (C.c_void_p * len(aps))(*[t.data_ptr() for t in dvec])

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def pipeline(timers=None):
        if timers is not None:
            timers[0].record(stream)
        mb._check(L.mvb200_super_batch(ctx.handle, C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), F))
        if timers is not None:
            timers[1].record(stream)
        # vectors and output only for this rank's own frames; the halo frames are references (their super frames above)
        mb._check(L.mvb200_analyse_multi_batch(ctx.handle, apa, len(aps), C.byref(sp), dsup.data_ptr(), OFF, F_out, outs, vs))
        if timers is not None:
            timers[2].record(stream)
        mb._check(L.mvb200_degrain_batch(ctx.handle, C.byref(dp), C.byref(sp), dsrc.data_ptr(), dsup.data_ptr(), F, OFF, F_out, outs, vs, ddst.data_ptr()))
        if timers is not None:
            timers[3].record(stream)

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

    # ---- device-resident timing
    for _ in range(args.warmup):
        pipeline()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    stage_events = [[ev() for _ in range(4)] for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        pipeline(stage_events[k])
    barrier()
    launches = ctx.launch_count() - launches0
    total_ms = sum(e[0].elapsed_time(e[3]) for e in stage_events)
    # the steps are back to back on one stream: whole-region time = first start -> last end
    region_ms = stage_events[0][0].elapsed_time(stage_events[-1][3])
    search_ms = sum(e[1].elapsed_time(e[2]) for e in stage_events) / args.steps
    super_ms = sum(e[0].elapsed_time(e[1]) for e in stage_events) / args.steps
    degrain_ms = sum(e[2].elapsed_time(e[3]) for e in stage_events) / args.steps
    region_ms = max_over_ranks(region_ms)
    value = world * F_out * args.steps / (region_ms / 1e3)

    # ---- end-to-end timing: pinned host source -> H2D -> pipeline -> D2H denoised frames
    # One C-ABI call per step with HOST buffers (mvb200_degrain_clip_host_async): H2D of the step's source
    # frames, Super + 6x Analyse + Degrain3, D2H of the denoised frames.  The call is asynchronous and the
    # engine double-buffers the clips, so the copies of neighbouring steps overlap the kernels; everything
    # (all copies included) completes inside the timed region.
    host_outs = [host_out, torch.zeros((F, cl.frame_bytes), dtype=torch.uint8).pin_memory()]
    e2e_calls = [0]

    def e2e_step():
        o = host_outs[e2e_calls[0] & 1]
        e2e_calls[0] += 1
        mb._check(L.mvb200_degrain_clip_host_async(ctx.handle, C.byref(sp), apa, len(aps), C.byref(dp), host.data_ptr(), o.data_ptr(), F, OFF, F_out))

    for _ in range(max(2, args.warmup)):
        e2e_step()
    barrier()
    e2e_launches0 = ctx.launch_count()
    t0 = time.perf_counter()
    e0, e1 = ev(), ev()
    e0.record(stream)
    for _ in range(args.steps):
        e2e_step()
    ctx.synchronize()  # compute stream and both copy streams of the engine
    e1.record(stream)
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * F_out * args.steps / (e2e_ms / 1e3)
    checksum = int(host_out.numpy()[:F_out, ::4099].astype(np.uint64).sum())
    if world > 1:  # the only collective of the job: reassemble the sharded clip's output planes (here: a digest of them)
        t = torch.tensor([checksum % (1 << 53)], dtype=torch.float64, device="cuda")
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_search): algorithmic bytes per launch
    bps = 1
    A = 1.5 * (WIDTH + 32) * (HEIGHT + 32) * bps
    S = (PEL * PEL + 1.0 / 3.0) * A          # reference side: the whole super frame
    S1 = (1 + 1.0 / 3.0) * A                  # source side: plane 0 of every level
    V = aps[0].vectors_size
    alg_bytes = len(aps) * F_out * (S1 + S + V)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (search_ms / 1e3) / 1e9
    # DRAM traffic of the same kernel from the committed ncu --set full capture (per chain, scaled to this launch)
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", NCU_PROFILE)))
        rd = prof["dram__bytes_read.sum"]; wr = prof["dram__bytes_write.sum"]
        tobytes = lambda m: m["value"] * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[m["unit"]]
        traffic = (tobytes(rd) + tobytes(wr)) / prof["_chains"] * len(aps) * F_out
        # the bound that actually binds a latency-serial integer search: issue slots / the integer ALU pipe (same capture)
        int_pipe = {"alu_pipe_active_pct_of_peak": prof["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]["value"],
                    "issue_active_pct_of_peak": prof["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"],
                    "warps_per_sm": prof["launch__occupancy_limit_registers"]["value"], "source": "profiles/" + NCU_PROFILE}
    except (OSError, KeyError, ValueError):
        int_pipe = None
    line = {
        "metric": "frames/sec Analyse+Degrain3 1080p YUV420P8", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": region_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": F_out, "halo_frames_per_gpu": F - F_out, "distinct_frames": DISTINCT,
                   "search_chains_per_step_per_gpu": len(aps) * F_out,
                   "l2": "inputs larger than L2 (%d MB of super frames + vectors per step vs 126 MB L2), no flush" % int(F * (sl.frame_bytes + 6 * vs) / 1e6),
                   "stage_ms": {"super": super_ms, "analyse_x6": search_ms, "degrain3": degrain_ms}},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(F * cl.frame_bytes), "d2h_bytes_per_step": int(F_out * cl.frame_bytes),
                "output_digest": checksum},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "k_search (6 analyse instances x F frames, one warp per search chain)", "bound": "hbm", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": ("dram__bytes_read+write of profiles/%s, per chain x chains of this launch" % NCU_PROFILE) if traffic else None,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "int_pipe": int_pipe,
                     "note": "the search is a serial dependency chain per (frame, direction, delta): issue/integer-pipe bound (see int_pipe), far below the HBM roofline by construction"},
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = DISTINCT  # one whole cycle of the synthetic clip, like the GPU arm's batch
        fps, err = reference_fps(sample, 4, cores)
        if fps is not None:
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                                    "sample": "%d frames of the same workload after 4 warm-up frames, one worker thread per host core, reference built from "
                                              "its own sources (x264 NASM kernels replaced by the reference's C fallbacks)" % sample}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": "unavailable: " + err}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.synchronize()
    torch.cuda.synchronize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=1184, help="frames per step per GPU (1184 x 6 chains = 4 full rounds of 12 warps x 148 SMs)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
