#!/bin/bash
# Round 2, GPU call h: long plugin runs (steady state without start / end effects), the SATD lane layout of the TMA kernel, c4.
cd "$(dirname "$0")/.."
TAG=${1:-r2h}
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
echo "== long plugin runs: 17760 frames (60 windows), 16 threads"
for c in 1 2 4; do
    r=$(MVB200_CONTEXTS_PER_DEVICE=$c timeout 300 python tools/bench_plugin.py --frames 17760 --threads 16 --lookahead auto --no-reference 2>&1 | tail -1 | cut -c1-160)
    echo "contexts=$c $r" | tee -a gpurun_out/${TAG}_long.txt
done
echo "== bench.py (default)"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print('value', round(d['value'],1), 'one-ctx', round(d['config']['value_one_context'],1), 'e2e', round(d['e2e']['value'],1), 'c_abi', round(d['e2e']['c_abi'],1), {k: round(v,1) for k,v in d['config']['stage_ms'].items()}, d['output_check']['mismatches'])"
echo "== c4 with the SATD lane layout"
timeout 500 python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_c4.json"))
    print("c4 value", round(d["value"], 1), "one-ctx", round(d["config"]["value_one_context"], 1), "e2e", round(d["e2e"]["value"], 1), d["config"]["search_kernel"], {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"])
except Exception as e:
    print("c4 FAILED", e); print(open("gpurun_out/${TAG}_bench_c4.err").read()[-1500:])
PY
du -sh gpurun_out
