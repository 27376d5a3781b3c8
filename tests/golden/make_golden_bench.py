"""Generates tests/golden/bench_digest.json: sha256 digests of the REFERENCE plugin's output (oracle/_ref/libmvtools_ref.so = the
reference sources compiled in place, driven by host/minivs) for every bench.py configuration AT FULL SIZE, on the bench's
own cyclic synthetic clip.  bench.py compares sampled output frames of its plugin leg with them (every rank its own shard),
tests/test_gpu_bench_workload.py the CUDA clip pipeline.  Runs in the build container only (needs /root/reference compiled).

    python tests/golden/make_golden_bench.py [config ...]

A clip of 2 * distinct + 8 frames is rendered: "head" = frames 0 .. 2 (no earlier frames exist), "cycle"[g % distinct] =
frames distinct .. 2 * distinct - 1 (an output frame depends on the frames within +-3 of it only, and frame i of the
clip is base[i % distinct], so every frame 3 <= g < N - 3 of a longer cyclic clip equals cycle[g % distinct])."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import minivs  # noqa: E402  (bench put host/minivs on the path)

OUT = os.path.join(HERE, "bench_digest.json")


def main():
    names = sys.argv[1:] or sorted(bench.CONFIGS)
    gold = json.load(open(OUT)) if os.path.exists(OUT) else {}
    threads = os.cpu_count() or 4
    for name in names:
        cfg = bench.CONFIGS[name]
        distinct = cfg.get("distinct", bench.DISTINCT)
        N = 2 * distinct + 8
        t0 = time.time()
        core = minivs.Core()
        mv = core.load(minivs.REF_PLUGIN)
        out, keep = bench.build_script(mv, core, cfg, bench.make_frames(cfg, N))
        keep[1].set_cache(16 + 2 * threads)
        for v in keep[2:]:
            v.set_cache(8 + 2 * threads)
        out.set_cache(N)
        out.render(0, N, threads)  # (parallel; the frames stay in the output node's cache)
        dig = [bench.frame_digest(cfg, out.get_frame(n)) for n in range(N)]
        for g in range(bench.REACH, distinct):  # the periodicity the golden relies on, checked on the reference itself
            assert dig[g] == dig[g + distinct] or g + distinct >= N - bench.REACH, (name, g)
        gold[name] = {"head": dig[:bench.REACH], "cycle": [dig[distinct + i] for i in range(distinct)], "clip_frames": N,
                      "workload": cfg["workload"]}
        json.dump(gold, open(OUT, "w"), indent=1, sort_keys=True)
        print(name, "done in %.1f s" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
