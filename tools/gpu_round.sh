#!/bin/bash
# One GPU-box call: parity tests, A/B bench of search-kernel variants, the ncu launch list of the
# bench command and one ncu --set full capture of the search kernel.  Everything lands in gpurun_out/
# (kept under 64 MiB or nothing comes back).
#   usage: tools/gpu_round.sh <tag> [variant[:frames] ...]   (variant = libmvtools_b200_<name>.so built by tools/build_variant.sh, or "default")
cd "$(dirname "$0")/.."
TAG=${1:-r1}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
echo "== pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
echo "== differential fuzz on the GPU: random scripts through the product plugin against the compiled reference plugin"
FUZZ_GLUE=gpu timeout 300 python tools/fuzz_oracle.py 1 60 > gpurun_out/${TAG}_fuzz_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_fuzz_gpu.log | cut -c1-300
MVB200_LOOKAHEAD=16 FUZZ_GLUE=gpu timeout 300 python tools/fuzz_oracle.py 2 60 > gpurun_out/${TAG}_fuzz_gpu_lookahead.log 2>&1; tail -3 gpurun_out/${TAG}_fuzz_gpu_lookahead.log | cut -c1-300
echo "== bench (b200 arm, default library)"
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
for spec in "$@"; do   # name[:frames]
    name=${spec%%:*}; frames=${spec#*:}; [ "$frames" = "$spec" ] && frames=1184
    lib=libmvtools_b200_$name.so; [ "$name" = default ] && lib=libmvtools_b200.so
    MVB200_LIB=$lib timeout 300 python bench.py --no-cpu-baseline --frames $frames > gpurun_out/${TAG}_bench_${name}_$frames.json 2> gpurun_out/${TAG}_bench_${name}_$frames.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_${name}_$frames.json"))
    print("variant $name F=$frames: fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["config"]["stage_ms"], "digest", d["e2e"]["output_digest"])
except Exception as e:
    print("variant $name F=$frames FAILED", e); print(open("gpurun_out/${TAG}_bench_${name}_$frames.err").read()[-600:])
PY
done
echo "== flow kernels (device-resident throughput; tools/bench_flow.py)"
timeout 300 python tools/bench_flow.py > gpurun_out/${TAG}_bench_flow.jsonl 2> gpurun_out/${TAG}_bench_flow.err
cat gpurun_out/${TAG}_bench_flow.jsonl | cut -c1-200; tail -2 gpurun_out/${TAG}_bench_flow.err
echo "== the plugin path: the canonical script through the mini host, look-ahead off / on (tools/bench_plugin.py)"
timeout 600 python tools/bench_plugin.py --frames 96 --threads 8 --lookahead 0,32,88 > gpurun_out/${TAG}_bench_plugin.jsonl 2> gpurun_out/${TAG}_bench_plugin.err
cat gpurun_out/${TAG}_bench_plugin.jsonl | cut -c1-200
echo "== bench (reference arm)"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
cat gpurun_out/${TAG}_bench_ref.json
echo "== ncu launch list of the bench command (not a bench value)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --frames 296 --no-cpu-baseline > gpurun_out/${TAG}_launches_run.log 2>&1
echo "launch list exit $?"
echo "== ncu --set full: search kernel (one launch)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_search_tma -c 1 -f -o gpurun_out/${TAG}_prof_search \
    python tools/gpu_prof.py ${PROF_FRAMES:-296} > gpurun_out/${TAG}_prof_search.log 2>&1
echo "search prof exit $?"
echo "== ncu --set full: the per-pixel flow kernels (one launch each; DESIGN.md 4.5 expects them near the HBM roofline)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_flow_fetch|k_flow_inter|k_flow_blur|k_blockfps|k_mask_plane|k_finest' -c 6 -f -o gpurun_out/${TAG}_prof_flow \
    python tools/bench_flow.py --frames 16 --reps 1 > gpurun_out/${TAG}_prof_flow.log 2>&1
echo "flow prof exit $?"
find gpurun_out -size +40M -print -delete
du -sh gpurun_out; ls -la gpurun_out | tail -25
