#!/bin/bash
# Round 2, GPU call e: contexts-per-device x window sweep through the plugin, the multi-context bench, Recalculate inside the fused chains.
cd "$(dirname "$0")/.."
TAG=${1:-r2e}
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
echo "== sweep: contexts per device x window (1080p Degrain3 script, 4440 frames, 16 host threads)"
for spec in 1:auto 2:auto 3:auto 4:auto 3:148 4:148 3:99 4:74 2:148 6:48; do
    c=${spec%%:*}; w=${spec#*:}
    if [ $c -le 4 ]; then export MVB200_CONTEXTS_PER_DEVICE=$c; unset MVB200_DEVICES; else export MVB200_DEVICES=0,0,0,0,0,0; fi
    r=$(timeout 200 python tools/bench_plugin.py --frames 4440 --threads 16 --lookahead $w --no-reference 2>&1 | tail -1 | cut -c1-160)
    echo "contexts=$c window=$w $r" | tee -a gpurun_out/${TAG}_sweep.txt
done
unset MVB200_CONTEXTS_PER_DEVICE MVB200_DEVICES
echo "== bench.py (default: 3 contexts)"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --no-cpu-baseline --contexts 4 --steps 2 > gpurun_out/${TAG}_bench_c4ctx.json 2> gpurun_out/${TAG}_bench_c4ctx.err
timeout 600 python bench.py --no-cpu-baseline --contexts 2 --steps 2 > gpurun_out/${TAG}_bench_c2ctx.json 2> gpurun_out/${TAG}_bench_c2ctx.err
python - <<PY
import json
for n in ("bench", "bench_c4ctx", "bench_c2ctx"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % n))
        print(n, "value", round(d["value"], 1), "one-ctx", round(d["config"]["value_one_context"], 1), "e2e", round(d["e2e"]["value"], 1), "c_abi", round(d["e2e"]["c_abi"] or 0, 1), {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"]["mismatches"])
    except Exception as e:
        print(n, "FAILED", e)
PY
echo "== c4 (Recalculate inside the fused Compensate chain)"
timeout 600 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_c4.json"))
    print("c4 value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"], d["e2e"]["protocol"][:60])
except Exception as e:
    print("c4 FAILED", e); print(open("gpurun_out/${TAG}_bench_c4.err").read()[-1500:])
PY
echo "== ncu: the fused Super kernel after the word-wise phase C"
timeout 200 ncu --set full --clock-control none -k regex:'k_super0_pel2' -c 1 -f -o gpurun_out/${TAG}_prof_super0_fused python tools/gpu_prof2.py 64 > gpurun_out/${TAG}_prof_super0_fused.log 2>&1
du -sh gpurun_out
