#!/bin/bash
# tests + smoke + bench on the GPU box; logs come back in gpurun_out/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^FAILED|^E   Assert" gpurun_out/pytest_gpu.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
