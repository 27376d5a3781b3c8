"""Round-2 measurement tool (not bench.py): frames/s of the canonical .vpy script
    sup = mv.Super(clip, pel=2); v.. = mv.Analyse(sup, blksize=16, overlap=8, delta=1..3, isb=0/1); mv.Degrain3(clip, sup, v..)
driven through a VapourSynth PLUGIN by the mini host, the way a frame server would (fmParallel, `threads` host threads):
this repo's plugin with MVB200_LOOKAHEAD off / on, and the compiled reference beside it.

    python tools/bench_plugin.py [--frames 96] [--size 1920x1080] [--threads 8] [--lookahead 0,32,96] [--plugin PATH]

Prints one JSON line per run.  --plugin tests/mock/libmvtools_mockengine.so exercises the tool itself on a CPU box."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("vapoursynth-mvtools_b200", os.path.join("host", "minivs")):
    sys.path.insert(0, os.path.join(ROOT, p))


def one_run(plugin, frames_n, w, h, threads, radius=3):
    """(runs in a child process: MVB200_LOOKAHEAD is read when the filters are created)"""
    import clipgen
    import minivs
    base = clipgen.make_clip(w, h, min(frames_n, 64), bits=8, seed=0)
    frames = [base[i % len(base)] for i in range(frames_n)]
    core = minivs.Core()
    mv = core.load(plugin)
    clip = core.source(frames, bits=8)
    sup = mv.Super(clip, pel=2)
    sup.set_cache(64 + 2 * threads)
    vecs = [mv.Analyse(sup, isb=isb, delta=d, blksize=16, overlap=8) for d in range(1, radius + 1) for isb in (1, 0)]
    for v in vecs:
        v.set_cache(8 + 2 * threads)
    out = mv.invoke("Degrain%d" % radius, clip, sup, *vecs)
    out.set_cache(0)
    warm = frames_n // 4
    out.render(0, warm, threads)
    t = out.render(warm, frames_n, threads)
    return (frames_n - warm) / t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=96)
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--lookahead", default="auto,32")
    ap.add_argument("--plugin", default=None)
    ap.add_argument("--child", default=None)
    ap.add_argument("--no-reference", action="store_true")
    a = ap.parse_args()
    w, h = (int(x) for x in a.size.split("x"))
    if a.child:
        print(json.dumps({"fps": one_run(a.child, a.frames, w, h, a.threads)}))
        return
    import minivs
    import mvtools_b200 as mb
    runs = [("reference", minivs.REF_PLUGIN, None)] if os.path.exists(minivs.REF_PLUGIN) and not a.no_reference else []
    for la in a.lookahead.split(","):  # "auto" = MVB200_LOOKAHEAD unset: the plugin's default
        runs.append(("b200 lookahead=%s" % la, a.plugin or mb.PLUGIN_PATH, la))
    for name, plugin, la in runs:
        env = dict(os.environ)
        env.pop("MVB200_LOOKAHEAD", None)
        if la and la != "auto":
            env["MVB200_LOOKAHEAD"] = la
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", plugin, "--frames", str(a.frames), "--size", a.size, "--threads", str(a.threads)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        try:
            fps = json.loads(r.stdout.strip().splitlines()[-1])["fps"]
            print(json.dumps({"run": name, "frames_per_s": fps, "frames": a.frames, "size": a.size, "threads": a.threads}), flush=True)
        except Exception:
            print(json.dumps({"run": name, "error": (r.stderr or r.stdout)[-400:]}), flush=True)


if __name__ == "__main__":
    main()
