#!/usr/bin/env python
"""The plugin's HOST plumbing alone (no GPU needed): the product's plugin.cpp over the stand-in engine of tests/mock with its compute
switched off (MVB200_MOCK_NOCOMPUTE), rendering a bench script -- frames requested, packed into windows, "computed", copied into
VapourSynth frames, properties attached.  Says how many frames/s the glue itself can carry on this machine's cores.
    python tools/bench_glue.py [--config c1] [--frames 4096] [--threads 8]"""
import argparse, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "host", "minivs"), os.path.join(ROOT, "vapoursynth-mvtools_b200"), os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c1")
    ap.add_argument("--frames", type=int, default=4096)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--compute", action="store_true")
    ap.add_argument("--latency-ms", type=int, default=1, help="what one (stand-in) engine call takes")
    a = ap.parse_args()
    if not a.compute:
        os.environ["MVB200_MOCK_NOCOMPUTE"] = str(max(1, a.latency_ms))
    import bench, minivs
    so = os.path.join(ROOT, "tests", "mock", "libmvtools_mockengine.so")
    cfg = bench.CONFIGS[a.config]
    core = minivs.Core()
    mv = core.load(so)
    out, keep = bench.build_script(mv, core, cfg, bench.make_frames(cfg, a.frames + 8 * int(os.environ.get("MVB200_LOOKAHEAD", "16") or 16), first=0))
    out.set_cache(0)
    warm = 6 * int(os.environ.get("MVB200_LOOKAHEAD", "16") or 16)
    out = out  # (warm-up: the windows' buffers get into circulation)
    out.render(0, warm, a.threads)
    t = out.render(warm, warm + a.frames, a.threads)
    print("%s: %d frames, %d threads: %.1f frames/s (%.3f s)" % (a.config, a.frames, a.threads, a.frames / t, t))


if __name__ == "__main__":
    main()
