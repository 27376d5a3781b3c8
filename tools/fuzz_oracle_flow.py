"""Differential fuzz of the flow-group oracle (oracle/mvtools_oracle_flow.c) against the compiled reference (oracle/_ref,
build container only): random clip formats, block shapes and filter arguments for mv.Mask / mv.Flow / mv.FlowInter /
mv.FlowFPS / mv.FlowBlur / mv.BlockFPS / mv.Finest / mv.SCDetection.  One configuration per process (the reference
corrupts its heap on some).  Test infrastructure (a tool, not part of the suite):

    python tools/fuzz_oracle_flow.py [seed] [iterations]"""
import os
import random
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one(seed):
    for p in ("vapoursynth-mvtools_b200", "oracle", os.path.join("host", "minivs"), "tests"):
        sys.path.insert(0, os.path.join(R, p))
    import pipeline
    import pipeline_flow
    import test_oracle_flow
    rnd = random.Random(seed)
    bw, bh = rnd.choice([(4, 4), (8, 4), (8, 8), (16, 8), (16, 16), (32, 16), (32, 32)])
    bits = rnd.choice([8, 8, 8, 10, 16])
    fmt = rnd.choice([dict(), dict(ssw=0, ssh=0), dict(ssw=1, ssh=0), dict(ssw=0, ssh=1), dict(gray=True)])
    ovx, ovy = rnd.choice([0, bw // 2]), rnd.choice([0, bh // 2])
    pel = rnd.choice([1, 2, 2, 4]) if bits != 8 else rnd.choice([1, 2, 2])  # (8-bit pel 4: the known last-column artefact shows in mv.Finest)
    W = rnd.choice([232, 264, 296]) + (8 if rnd.random() < 0.5 else 0)
    H = rnd.choice([136, 160, 200])
    while any((-(w + 16) * ((bits + 7) // 8)) % 64 < 16 for w in (W, W // 2)):
        W += 8
    ck = dict(width=W, height=H, nframes=5, bits=bits, **fmt)
    if rnd.random() < 0.3:
        ck["scene_cut"] = True
    skw = dict(pel=pel, hpad=8, vpad=8)
    akw = dict(blksize=bw, blksizev=bh, overlap=ovx, overlapv=ovy, search=rnd.choice([3, 4, 5]), searchparam=rnd.choice([1, 2]))
    f = lambda a, b: round(rnd.uniform(a, b), 1)
    flow = dict(
        flow=[dict(time=f(0, 100), mode=rnd.choice([0, 1]), vec=rnd.choice([0, 1]))],
        scd=[dict(vec=rnd.choice([0, 1]), thscd1=rnd.choice([100, 400, 2000]), thscd2=rnd.choice([30, 130, 255]))],
        # (the reference's mv.Finest overruns subsampled chroma planes at pel 1, MVFinest.c:75-77)
        finest=pel > 1 or fmt.get("gray", False) or (fmt.get("ssw", 1) == 0 and fmt.get("ssh", 1) == 0),
        inter=[dict(time=f(1, 99), ml=f(5, 200), blend=rnd.choice([0, 1]))],
        fps=[dict(num=rnd.choice([30, 50, 60, 75]), den=rnd.choice([1, 2]), mask=rnd.choice([0, 1, 2]), ml=f(5, 200), blend=rnd.choice([0, 1]))],
        blur=[dict(blur=f(1, 200), prec=rnd.choice([1, 2, 3]))],
        bfps=[dict(num=rnd.choice([30, 50, 60]), den=1, mode=rnd.choice(range(9)), ml=f(5, 200), blend=rnd.choice([0, 1]))])
    if bits == 8:
        flow["mask"] = [dict(kind=rnd.choice(range(6)), ml=f(5, 200), gamma=f(0.3, 3.0), time=f(0, 100), vec=rnd.choice([0, 1]), ysc=rnd.choice([0, 99, 255]))]
    name = "fuzz"
    pipeline.CASES[name] = (ck, skw, akw, 1, {})
    pipeline_flow.FLOW[name] = flow
    ref = pipeline_flow.run_reference_flow(name)
    got = test_oracle_flow.oracle_run(name)
    bad = [st for st in ("scd", "mask", "flow", "finest", "inter", "fps", "blur", "bfps") if got.get(st, []) != ref.get(st, [])]
    print("seed %d: %s %s" % (seed, "ok" if not bad else "MISMATCH %s" % bad, "" if not bad else (ck, skw, akw, flow)), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        one(int(sys.argv[2]))
        sys.exit(0)
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    bad = 0
    for s in range(seed0, seed0 + (int(sys.argv[2]) if len(sys.argv) > 2 else 20)):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(s)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        line = [x for x in r.stdout.splitlines() if x.startswith("seed")]
        if r.returncode != 0 or not line:
            print("seed %d: the run died (exit %d): %s" % (s, r.returncode, r.stdout.strip().splitlines()[-1][:200] if r.stdout.strip() else ""))
        else:
            print(line[0])
            bad += "MISMATCH" in line[0]
    print("done, mismatches =", bad)
