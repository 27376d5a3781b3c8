#!/bin/bash
# Round 2, GPU call l: where the plugin's host threads spend their time on configs[0] (MVB200_STATS), c4 with a warm-up that
# covers the pinned buffers' circulation.
cd "$(dirname "$0")/.."
TAG=${1:-r2l}
mkdir -p gpurun_out
for c in c1 c4; do
    MVB200_STATS=1 timeout 600 python bench.py --config $c --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
    echo "config $c exit $?"
    grep 'mvtools-b200' gpurun_out/${TAG}_bench_$c.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$c.json"))
    print("$c value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["output_check"]["mismatches"], "cpu", (d.get("cpu_baseline") or {}).get("value"), d["e2e"].get("protocol"))
except Exception as e:
    print("$c FAILED", e); print(open("gpurun_out/${TAG}_bench_$c.err").read()[-1200:])
PY
done
nproc; free -g | head -2
