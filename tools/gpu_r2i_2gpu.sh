#!/bin/bash
# Round 2, 2-GPU call: the two-rank test, the plugin's own multi-GPU mode (one process, windows round every device), the 2-rank bench.
cd "$(dirname "$0")/.."
TAG=${1:-r2i}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt; nproc >> gpurun_out/${TAG}_gpus.txt; free -g | head -2 >> gpurun_out/${TAG}_gpus.txt; cat gpurun_out/${TAG}_gpus.txt
echo "== two-rank test"
timeout 600 python -m pytest tests/test_gpu_bench_workload.py -m gpu -q -k "two_rank" > gpurun_out/${TAG}_pytest_2rank.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_2rank.log
echo "== the plugin on one GPU and on both (one process, 17760 frames)"
for dev in 0 "0,1"; do
    r=$(MVB200_DEVICES=$dev timeout 300 python tools/bench_plugin.py --frames 17760 --threads $(nproc) --lookahead auto --no-reference 2>&1 | tail -1 | cut -c1-160)
    echo "MVB200_DEVICES=$dev $r" | tee -a gpurun_out/${TAG}_plugin_multigpu.txt
done
r=$(timeout 300 python tools/bench_plugin.py --frames 17760 --threads $(nproc) --lookahead auto --no-reference 2>&1 | tail -1 | cut -c1-160)
echo "default (all devices) $r" | tee -a gpurun_out/${TAG}_plugin_multigpu.txt
echo "== bench.py under torchrun, 2 ranks"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench_2gpu.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_bench_2gpu.json") if l.startswith("{")][-1])
    print("N=2 value", round(d["value"], 1), "one-ctx", round(d["config"]["value_one_context"], 1), "e2e", round(d["e2e"]["value"], 1), "c_abi", round(d["e2e"]["c_abi"], 1), d["output_check"])
except Exception as e:
    print("FAILED", e)
PY
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_1gpu.json')); print('N=1 (same box) value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'c_abi', round(d['e2e']['c_abi'],1))"
du -sh gpurun_out
