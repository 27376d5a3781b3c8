#!/bin/bash
# Round 2, GPU call j: row-parallel SATD in the generic kernels (parity + c4), every configuration with the final protocol,
# one ncu capture of the flow group's kernels, the final headline line.
cd "$(dirname "$0")/.."
TAG=${1:-r2j}
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== configurations"
for c in c4 c2 c3 c5 c1; do
    timeout 600 python bench.py --config $c --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
    echo "config $c exit $?"
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$c.json"))
    print("$c value", round(d["value"], 1), "one-ctx", round(d["config"]["value_one_context"], 1), "e2e", round(d["e2e"]["value"], 1), "c_abi", d["e2e"].get("c_abi"), d["config"]["search_kernel"][:14], {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"]["mismatches"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$c FAILED", e); print(open("gpurun_out/${TAG}_bench_$c.err").read()[-1200:])
PY
done
echo "== ncu --set full: flow group kernels (one launch each)"
timeout 300 ncu --set full --clock-control none -k regex:'k_flow_fetch|k_flow_inter|k_flow_blur|k_mask_plane|k_finest|k_small_vectors' -c 6 -f -o gpurun_out/${TAG}_prof_flow \
    python tools/bench_flow.py --frames 16 --reps 1 > gpurun_out/${TAG}_prof_flow.log 2>&1
echo "flow prof exit $?"
echo "== bench.py (default, with cpu_baseline)"
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print('value', round(d['value'],1), 'one-ctx', round(d['config']['value_one_context'],1), 'e2e', round(d['e2e']['value'],1), 'c_abi', round(d['e2e']['c_abi'],1), {k: round(v,1) for k,v in d['config']['stage_ms'].items()}, d['output_check']['mismatches'], d['cpu_baseline']['value'], d['clocks'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
find gpurun_out -name '*.ncu-rep' -size +30M -print -delete
du -sh gpurun_out
