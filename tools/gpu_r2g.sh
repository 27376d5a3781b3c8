#!/bin/bash
# Round 2, GPU call g: hunt the one intermittent Compensate mismatch; parallel packing; de-phased resident leg; launch list.
cd "$(dirname "$0")/.."
TAG=${1:-r2g}
mkdir -p gpurun_out
PY=$(which python)
echo "== stress: per-frame Compensate with plugin windows in flight"
timeout 300 python tools/stress_compensate.py 40 > gpurun_out/${TAG}_stress.log 2>&1; tail -5 gpurun_out/${TAG}_stress.log
echo "== initcheck of the per-frame path of that case"
timeout 400 compute-sanitizer --tool initcheck $PY -m pytest tests/test_zz_gpu_pending.py -m gpu -x -q -k "test_cuda_matches_reference_goldens_late_cases and blk128x64_gray_p8" > gpurun_out/${TAG}_initcheck.log 2>&1
grep -c "Uninitialized" gpurun_out/${TAG}_initcheck.log; grep -B2 -A14 "Uninitialized" gpurun_out/${TAG}_initcheck.log | head -60; tail -3 gpurun_out/${TAG}_initcheck.log
echo "== pytest -m gpu (twice)"
for i in 1 2; do timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu_$i.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/${TAG}_pytest_gpu_$i.log; done
echo "== bench.py (default) x2"
for i in 1 2; do
    timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_$i.json 2> gpurun_out/${TAG}_bench_$i.err
    python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_$i.json')); print('value', round(d['value'],1), 'one-ctx', round(d['config']['value_one_context'],1), 'e2e', round(d['e2e']['value'],1), 'c_abi', round(d['e2e']['c_abi'],1), {k: round(v,1) for k,v in d['config']['stage_ms'].items()}, d['output_check']['mismatches'])"
done
echo "== plugin sweep with parallel packing"
for c in 2 4 6; do
    r=$(MVB200_CONTEXTS_PER_DEVICE=$c timeout 200 python tools/bench_plugin.py --frames 5920 --threads 16 --lookahead auto --no-reference 2>&1 | tail -1 | cut -c1-160)
    echo "contexts=$c window=auto $r" | tee -a gpurun_out/${TAG}_sweep.txt
done
echo "== ncu launch list of the bench command (not a bench value)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --frames 296 --no-cpu-baseline > gpurun_out/${TAG}_launches_run.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/${TAG}_launches.csv
echo "== c2 / c3"
for c in c2 c3; do
    timeout 500 python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$c.json"))
    print("$c value", round(d["value"], 1), "one-ctx", round(d["config"]["value_one_context"], 1), "e2e", round(d["e2e"]["value"], 1), "c_abi", d["e2e"].get("c_abi"), {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"]["mismatches"])
except Exception as e:
    print("$c FAILED", e); print(open("gpurun_out/${TAG}_bench_$c.err").read()[-1200:])
PY
done
du -sh gpurun_out
