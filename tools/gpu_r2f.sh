#!/bin/bash
# Round 2, GPU call f: the one failing GPU test under compute-sanitizer, the window-stream bench, more contexts, all configs.
cd "$(dirname "$0")/.."
TAG=${1:-r2f}
mkdir -p gpurun_out
echo "== the late Compensate case alone, then under compute-sanitizer (initcheck, memcheck)"
timeout 300 python -m pytest tests/test_zz_gpu_pending.py -m gpu -x -q -k "late_cases and blk128x64_gray_p8" > gpurun_out/${TAG}_late_alone.log 2>&1; tail -2 gpurun_out/${TAG}_late_alone.log
timeout 600 compute-sanitizer --tool initcheck --track-unused-memory no python -m pytest tests/test_zz_gpu_pending.py -m gpu -x -q -k "test_cuda_matches_reference_goldens_late_cases and blk128x64_gray_p8" > gpurun_out/${TAG}_initcheck.log 2>&1
grep -c "Uninitialized" gpurun_out/${TAG}_initcheck.log; grep -A12 "Uninitialized" gpurun_out/${TAG}_initcheck.log | head -40; tail -3 gpurun_out/${TAG}_initcheck.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_zz_gpu_pending.py -m gpu -x -q -k "test_cuda_matches_reference_goldens_late_cases and blk128x64_gray_p8" > gpurun_out/${TAG}_memcheck.log 2>&1
grep -A12 "Invalid" gpurun_out/${TAG}_memcheck.log | head -30; tail -3 gpurun_out/${TAG}_memcheck.log
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
echo "== bench.py (default: windows streamed through 4 contexts)"
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("value", round(d["value"], 1), "one-ctx", round(d["config"]["value_one_context"], 1), "e2e", round(d["e2e"]["value"], 1), "c_abi", round(d["e2e"]["c_abi"] or 0, 1), {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"]["mismatches"], d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("FAILED", e)
PY
echo "== more contexts"
for c in 4 6 8; do
    r=$(MVB200_CONTEXTS_PER_DEVICE=$c timeout 200 python tools/bench_plugin.py --frames 5920 --threads 16 --lookahead auto --no-reference 2>&1 | tail -1 | cut -c1-160)
    echo "contexts=$c window=auto $r" | tee -a gpurun_out/${TAG}_sweep.txt
done
for c in 6 8; do
    timeout 300 python bench.py --no-cpu-baseline --contexts $c --steps 2 > gpurun_out/${TAG}_bench_ctx$c.json 2> gpurun_out/${TAG}_bench_ctx$c.err
    python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_ctx$c.json')); print('resident contexts=$c value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
done
echo "== configurations"
for c in c2 c3 c5 c4 c1; do
    timeout 500 python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
    echo "config $c exit $?"
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$c.json"))
    print("$c value", round(d["value"], 1), "one-ctx", round(d["config"]["value_one_context"], 1), "e2e", round(d["e2e"]["value"], 1), "c_abi", d["e2e"].get("c_abi"), {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"]["mismatches"], d["e2e"]["protocol"][:30])
except Exception as e:
    print("$c FAILED", e); print(open("gpurun_out/${TAG}_bench_$c.err").read()[-1200:])
PY
done
du -sh gpurun_out
