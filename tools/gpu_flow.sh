#!/bin/bash
# A short GPU-box call: parity tests of the SURVEY 8f kernels (flow.cu) only.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout ${PYTEST_TIMEOUT:-130} python -m pytest tests/test_gpu_flow.py tests/test_plugin.py -m gpu -q -k "flow" > gpurun_out/${1:-flow}_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${1:-flow}_pytest.log
tail -40 gpurun_out/${1:-flow}_pytest.log | cut -c1-300
