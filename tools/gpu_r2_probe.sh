#!/bin/bash
# Round 2, first GPU-box call: the evidence the round's work is planned on (nothing here is a bench value).
#   1. ncu --set full --import-source on of the search kernel, one full wave of the bench clip (.ncu-rep comes back)
#   2. the same for the Super and Degrain kernels (16 frames)
#   3. the plugin path through the mini host as it stood at the end of round 1 (look-ahead off / 32 / 256)
#   4. the differential fuzz on the GPU (product plugin + CUDA engine against the compiled reference plugin)
#   5. flow-group throughput (tools/bench_flow.py)
#   6. the search's tail: stage times at 296 and 1184 frames per step
cd "$(dirname "$0")/.."
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc > gpurun_out/${TAG}_nproc.txt; free -g >> gpurun_out/${TAG}_nproc.txt
echo "== ncu --set full: search kernel, 1776 chains"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_search_tma -c 1 -f -o gpurun_out/${TAG}_prof_search \
    python tools/gpu_prof.py 296 > gpurun_out/${TAG}_prof_search.log 2>&1
echo "search prof exit $?"
echo "== ncu --set full: Super + Degrain kernels, 16 frames"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_fill|k_pad|k_reduce|k_refine|k_avg2|k_scd|k_weights|k_blend' -c 40 -f -o gpurun_out/${TAG}_prof_sd \
    python tools/gpu_prof2.py 16 > gpurun_out/${TAG}_prof_sd.log 2>&1
echo "super/degrain prof exit $?"
echo "== plugin path (round-1 glue)"
timeout 420 python tools/bench_plugin.py --frames 320 --threads 16 --lookahead 0,32,256 > gpurun_out/${TAG}_bench_plugin.jsonl 2> gpurun_out/${TAG}_bench_plugin.err
cat gpurun_out/${TAG}_bench_plugin.jsonl | cut -c1-240
echo "== stage times at 296 / 1184 frames per step"
for f in 296 1184; do
    timeout 200 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --frames $f > gpurun_out/${TAG}_bench_f$f.json 2> gpurun_out/${TAG}_bench_f$f.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_f$f.json"))
    print("F=$f fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["config"]["stage_ms"])
except Exception as e:
    print("F=$f FAILED", e); print(open("gpurun_out/${TAG}_bench_f$f.err").read()[-600:])
PY
done
echo "== flow kernels"
timeout 240 python tools/bench_flow.py > gpurun_out/${TAG}_bench_flow.jsonl 2> gpurun_out/${TAG}_bench_flow.err
cat gpurun_out/${TAG}_bench_flow.jsonl | cut -c1-220; tail -2 gpurun_out/${TAG}_bench_flow.err
echo "== differential fuzz on the GPU"
FUZZ_GLUE=gpu timeout 200 python tools/fuzz_oracle.py 1 40 > gpurun_out/${TAG}_fuzz_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_fuzz_gpu.log | cut -c1-300
MVB200_LOOKAHEAD=16 FUZZ_GLUE=gpu timeout 200 python tools/fuzz_oracle.py 2 40 > gpurun_out/${TAG}_fuzz_gpu_lookahead.log 2>&1; tail -3 gpurun_out/${TAG}_fuzz_gpu_lookahead.log | cut -c1-300
find gpurun_out -size +45M -print -delete
du -sh gpurun_out; ls -la gpurun_out | tail -25
