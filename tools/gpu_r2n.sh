#!/bin/bash
# Round 2, GPU call n: every fused clip call on the shared asynchronous pipeline (copies of neighbouring windows beside the
# kernels), best-fit pinned pool, adaptive warm-up of the plugin leg: parity, then c1 / c4 / c5 / the default line.
cd "$(dirname "$0")/.."
TAG=${1:-r2n}
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== configurations"
for c in c1 c4 c5 c3; do
    MVB200_STATS=1 timeout 600 python bench.py --config $c --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
    echo "config $c exit $?"
    grep 'mvtools-b200' gpurun_out/${TAG}_bench_$c.err | cut -c1-330
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$c.json"))
    print("$c value", round(d["value"], 1), "one-ctx", round(d["config"]["value_one_context"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"]["mismatches"], "cpu", (d.get("cpu_baseline") or {}).get("value"), d["e2e"].get("protocol"))
except Exception as e:
    print("$c FAILED", e); print(open("gpurun_out/${TAG}_bench_$c.err").read()[-1200:])
PY
done
echo "== bench.py (default, with cpu_baseline)"
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print('value', round(d['value'],1), 'one-ctx', round(d['config']['value_one_context'],1), 'e2e', round(d['e2e']['value'],1), 'c_abi', round(d['e2e']['c_abi'],1), {k: round(v,1) for k,v in d['config']['stage_ms'].items()}, d['output_check']['mismatches'], d['cpu_baseline']['value'], d['clocks'], d['e2e']['protocol'])"
du -sh gpurun_out
