#!/bin/bash
# Round 2, GPU call c: ncu captures of the Degrain and Super kernels at a realistic batch (64 frames), the new GPU tests,
# and a first pass over every BASELINE configuration (bench.py --config c1..c5).
cd "$(dirname "$0")/.."
TAG=${1:-r2c}
mkdir -p gpurun_out
echo "== new GPU tests"
timeout 600 python -m pytest tests/test_gpu_bench_workload.py tests/test_zz_gpu_pending.py -m gpu -x -q > gpurun_out/${TAG}_pytest_new.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest_new.log
echo "== ncu: Degrain kernels (64 frames)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_scd|k_weights|k_blend' -c 8 -f -o gpurun_out/${TAG}_prof_degrain \
    python tools/gpu_prof2.py 64 > gpurun_out/${TAG}_prof_degrain.log 2>&1
echo "degrain prof exit $?"
echo "== ncu: Super level-0 kernels (64 frames)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_fill_w|k_pad_w|k_refine_hv_w' -c 6 -f -o gpurun_out/${TAG}_prof_super0 \
    python tools/gpu_prof2.py 64 > gpurun_out/${TAG}_prof_super0.log 2>&1
echo "super prof exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_reduce|k_pad' -c 4 -f -o gpurun_out/${TAG}_prof_super1 \
    python tools/gpu_prof2.py 64 > gpurun_out/${TAG}_prof_super1.log 2>&1
echo "== every BASELINE configuration, one step each"
for c in c1 c2 c3 c5 c4; do
    timeout 420 python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
    echo "config $c exit $?"
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$c.json"))
    print("$c value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "c_abi", d["e2e"].get("c_abi"), d["config"]["search_kernel"], {k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["output_check"])
except Exception as e:
    print("$c FAILED", e); print(open("gpurun_out/${TAG}_bench_$c.err").read()[-1200:])
PY
done
find gpurun_out -name '*.ncu-rep' -size +30M -print -delete
du -sh gpurun_out
