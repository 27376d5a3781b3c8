"""Generates tests/golden/golden.json by running the UNMODIFIED reference
(oracle/_ref/libmvtools_ref.so, built in place from /root/reference by oracle/Makefile)
through the mini VapourSynth host on the seeded clips of tests/pipeline.py.
Run from the repo root in the build container:  python tests/golden/make_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in ("vapoursynth-mvtools_b200", "oracle", os.path.join("host", "minivs"), "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import pipeline  # noqa: E402

out = {}
for case in pipeline.CASES:
    out[case] = pipeline.run_reference(case)
    print(case, "ok", flush=True)
import minivs  # noqa: E402
reg = minivs.Core().load(minivs.REF_PLUGIN)
with open(os.path.join(HERE, "registration.json"), "w") as f:
    json.dump({"identifier": reg.identifier, "namespace": reg.namespace, "functions": reg.functions()}, f, indent=0, sort_keys=True)
with open(os.path.join(HERE, "golden.json"), "w") as f:
    json.dump(out, f, indent=0, sort_keys=True)
print("wrote golden.json")
