#!/bin/bash
# Full GPU test pass with logs brought back in gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -40 gpurun_out/pytest_gpu.log
