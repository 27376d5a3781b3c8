#!/bin/bash
# Round 2, GPU call t: final numbers -- c4 / c5 / c2 / c3 with the resident leg's windows going round every context, the default
# line with cpu_baseline, the reference arm, smoke.
cd "$(dirname "$0")/.."
TAG=${1:-r2t}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for c in c4 c5 c2 c3; do
    timeout 600 python bench.py --config $c --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
    echo "config $c exit $?"
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$c.json"))
    print("$c value", round(d["value"], 1), "one-ctx", round(d["config"]["value_one_context"], 1), "e2e", round(d["e2e"]["value"], 1), d["config"]["contexts_per_gpu"], d["config"]["window_frames"], {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"]["mismatches"], "cpu", (d.get("cpu_baseline") or {}).get("value"), d["e2e"]["protocol"][:60])
except Exception as e:
    print("$c FAILED", e); print(open("gpurun_out/${TAG}_bench_$c.err").read()[-1500:])
PY
done
echo "== bench.py (default, with cpu_baseline)"
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print('value', round(d['value'],1), 'one-ctx', round(d['config']['value_one_context'],1), 'e2e', round(d['e2e']['value'],1), 'c_abi', round(d['e2e']['c_abi'],1), {k: round(v,1) for k,v in d['config']['stage_ms'].items()}, d['output_check']['mismatches'], d['cpu_baseline']['value'], d['clocks'], d['roofline'], d['gpu_launches'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
