#!/bin/bash
# Round 2, GPU call p: the table-upload race -- reproducer (library built with the old synchronous upload) against the fix, then
# the GPU suite twice.
cd "$(dirname "$0")/.."
TAG=${1:-r2p}
mkdir -p gpurun_out
MVB200_LIB=libmvtools_b200_synctab.so timeout 300 python tools/stress_tables.py 400 2>&1 | tail -2 | tee gpurun_out/${TAG}_stress_tables.txt
timeout 300 python tools/stress_tables.py 400 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_stress_tables.txt
for i in 1 2; do
    timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu_$i.log 2>&1
    echo "pytest $i exit $?"; tail -2 gpurun_out/${TAG}_pytest_gpu_$i.log
done
