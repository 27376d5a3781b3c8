#!/bin/bash
# Round 2, GPU call r: batched Compensate (multi-CTA scene-change count, grid.z = frame): parity, c4; the differential fuzz on the
# final library (new seeds; default windows and 16-frame windows).
cd "$(dirname "$0")/.."
TAG=${1:-r2r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/${TAG}_pytest_gpu.log
for c in c4; do
    timeout 600 python bench.py --config $c --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
    echo "config $c exit $?"
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$c.json"))
    print("$c value", round(d["value"], 1), "one-ctx", round(d["config"]["value_one_context"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"]["mismatches"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$c FAILED", e); print(open("gpurun_out/${TAG}_bench_$c.err").read()[-1200:])
PY
done
FUZZ_GLUE=gpu timeout 240 python tools/fuzz_oracle.py 11 50 > gpurun_out/${TAG}_fuzz_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_fuzz_gpu.log | cut -c1-300
MVB200_LOOKAHEAD=16 FUZZ_GLUE=gpu timeout 240 python tools/fuzz_oracle.py 12 50 > gpurun_out/${TAG}_fuzz_gpu_lookahead.log 2>&1; tail -2 gpurun_out/${TAG}_fuzz_gpu_lookahead.log | cut -c1-300
