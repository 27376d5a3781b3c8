#!/bin/bash
# Round 2, GPU call s: the window plan chooses the engine contexts (more, smaller windows where a window cannot be a whole wave of
# chains: 4K); parity, then c4 / c5 / c2 / c1.
cd "$(dirname "$0")/.."
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
for c in c4 c5 c2 c1; do
    MVB200_STATS=1 timeout 600 python bench.py --config $c --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
    echo "config $c exit $?"
    grep 'mvtools-b200' gpurun_out/${TAG}_bench_$c.err | cut -c1-300
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$c.json"))
    print("$c value", round(d["value"], 1), "one-ctx", round(d["config"]["value_one_context"], 1), "e2e", round(d["e2e"]["value"], 1), d["config"]["contexts_per_gpu"], d["config"]["window_frames"], {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"]["mismatches"], "cpu", (d.get("cpu_baseline") or {}).get("value"), d["e2e"]["protocol"][:60])
except Exception as e:
    print("$c FAILED", e); print(open("gpurun_out/${TAG}_bench_$c.err").read()[-1500:])
PY
done
