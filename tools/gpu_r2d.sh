#!/bin/bash
# Round 2, GPU call d: the fused level-0 Super kernel (parity + A/B), two engine contexts on one device (overlap of
# consecutive windows), the fixed bench legs, every configuration again.
cd "$(dirname "$0")/.."
TAG=${1:-r2d}
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
echo "== bench.py (default) and the separate-pass Super for comparison"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
MVB200_NO_FUSED_SUPER=1 timeout 600 python bench.py --no-cpu-baseline --steps 2 > gpurun_out/${TAG}_bench_nofused.json 2> gpurun_out/${TAG}_bench_nofused.err
python - <<PY
import json
for n in ("bench", "bench_nofused"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % n))
        print(n, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "c_abi", round(d["e2e"]["c_abi"] or 0, 1), {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"]["mismatches"])
    except Exception as e:
        print(n, "FAILED", e)
PY
echo "== plugin: one context per device vs two contexts on the device"
timeout 300 python tools/bench_plugin.py --frames 3552 --threads 16 --lookahead auto --no-reference > gpurun_out/${TAG}_plugin_1ctx.jsonl 2>&1; cat gpurun_out/${TAG}_plugin_1ctx.jsonl | cut -c1-200
MVB200_DEVICES=0,0 timeout 300 python tools/bench_plugin.py --frames 3552 --threads 16 --lookahead auto --no-reference > gpurun_out/${TAG}_plugin_2ctx.jsonl 2>&1; cat gpurun_out/${TAG}_plugin_2ctx.jsonl | cut -c1-200
MVB200_DEVICES=0,0,0 timeout 300 python tools/bench_plugin.py --frames 3552 --threads 16 --lookahead auto --no-reference > gpurun_out/${TAG}_plugin_3ctx.jsonl 2>&1; cat gpurun_out/${TAG}_plugin_3ctx.jsonl | cut -c1-200
MVB200_DEVICES=0,0 timeout 300 python tools/bench_plugin.py --frames 3552 --threads 16 --lookahead 148 --no-reference > gpurun_out/${TAG}_plugin_2ctx_148.jsonl 2>&1; cat gpurun_out/${TAG}_plugin_2ctx_148.jsonl | cut -c1-200
echo "== ncu --set full: the fused Super kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_super0_pel2' -c 2 -f -o gpurun_out/${TAG}_prof_super0_fused \
    python tools/gpu_prof2.py 64 > gpurun_out/${TAG}_prof_super0_fused.log 2>&1
echo "fused super prof exit $?"
echo "== configurations"
for c in c2 c3 c5 c1; do
    timeout 420 python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
    echo "config $c exit $?"
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$c.json"))
    print("$c value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "c_abi", d["e2e"].get("c_abi"), d["config"]["search_kernel"], {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"], d["e2e"]["protocol"][:40])
except Exception as e:
    print("$c FAILED", e); print(open("gpurun_out/${TAG}_bench_$c.err").read()[-1200:])
PY
done
find gpurun_out -name '*.ncu-rep' -size +30M -print -delete
du -sh gpurun_out
