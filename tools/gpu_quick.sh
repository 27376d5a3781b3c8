#!/bin/bash
# A short GPU-box call (a few minutes): parity of the production search kernel against the goldens,
# then an A/B bench of the default library against named variants.  Results land in gpurun_out/ as they
# are produced, so a call cut off by the budget still returns what finished.
#   usage: tools/gpu_quick.sh <tag> [variant ...]
cd "$(dirname "$0")/.."
TAG=${1:-quick}; shift
mkdir -p gpurun_out
echo "== parity (goldens + oracle bytes + batch path)"
timeout ${PYTEST_TIMEOUT:-170} python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "test_cuda_matches_reference_goldens or test_cuda_matches_oracle_bytes or test_batch_path_matches_goldens" \
    > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_pytest_gpu.log
for name in default "$@"; do
    lib=libmvtools_b200_$name.so; [ "$name" = default ] && lib=libmvtools_b200.so
    MVB200_LIB=$lib timeout 120 python bench.py --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
    echo "bench $name exit $?"; cat gpurun_out/${TAG}_bench_${name}.json | cut -c1-600
done
