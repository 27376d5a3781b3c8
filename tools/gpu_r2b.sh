#!/bin/bash
# Round 2, GPU call b: the new plugin path (aligned / pipelined look-ahead windows, on by default) -- parity suite, the bench
# workload pinned at full size, the plugin's throughput through the mini host, and the new bench.py line.
cd "$(dirname "$0")/.."
TAG=${1:-r2b}
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
echo "== plugin path through the mini host"
timeout 400 python tools/bench_plugin.py --frames 1480 --threads 16 --lookahead auto,148 > gpurun_out/${TAG}_bench_plugin.jsonl 2> gpurun_out/${TAG}_bench_plugin.err
cat gpurun_out/${TAG}_bench_plugin.jsonl | cut -c1-240; tail -3 gpurun_out/${TAG}_bench_plugin.err
echo "== bench.py (default)"
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
echo "== bench.py --impl reference"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
cat gpurun_out/${TAG}_bench_ref.json | cut -c1-400
du -sh gpurun_out
