#!/bin/bash
# usage: tools/variant_bench.sh "<nvcc -D flags>" <frames> <tag>   -- rebuilds the TMA search kernel with the flags and runs the bench (GPU box)
set -e
cd "$(dirname "$0")/../vapoursynth-mvtools_b200"
rm -f obj/search_tma.o
make EXTRA="$1" libmvtools_b200.so > /dev/null
grep -E "spill|registers" obj/search_tma.ptxas.log | tail -2
cd ..
python bench.py --no-cpu-baseline --frames "$2" --steps 3 --warmup 3 > gpurun_out/bench_$3.json 2> gpurun_out/bench_$3.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$3.json"))
print("$3", "fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["config"]["stage_ms"])
PY
