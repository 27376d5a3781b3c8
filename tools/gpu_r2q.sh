#!/bin/bash
# Round 2, GPU call q: Recalculate on the chain's candidate layout (k_recalc_tma): parity, c4.
cd "$(dirname "$0")/.."
TAG=${1:-r2q}
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
