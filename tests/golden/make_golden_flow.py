"""Generates tests/golden/golden_flow.json: mv.SCDetection / mv.Mask / mv.Finest / mv.Flow of the UNMODIFIED reference
(oracle/_ref/libmvtools_ref.so, built in place from /root/reference by oracle/Makefile) driven by the mini VapourSynth
host on the seeded clips of tests/pipeline_flow.py.  Also refreshes registration.json (all registered functions).
Run from the repo root in the build container:  python tests/golden/make_golden_flow.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in ("vapoursynth-mvtools_b200", "oracle", os.path.join("host", "minivs"), "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import pipeline_flow  # noqa: E402

out = {}
for case in pipeline_flow.FLOW:
    out[case] = pipeline_flow.run_reference_flow(case)
    print(case, "ok", flush=True)
with open(os.path.join(HERE, "golden_flow.json"), "w") as f:
    json.dump(out, f, indent=0, sort_keys=True)
import minivs  # noqa: E402
reg = minivs.Core().load(minivs.REF_PLUGIN)
with open(os.path.join(HERE, "registration.json"), "w") as f:
    json.dump({"identifier": reg.identifier, "namespace": reg.namespace, "functions": reg.functions()}, f, indent=0, sort_keys=True)
print("wrote golden_flow.json, registration.json")
