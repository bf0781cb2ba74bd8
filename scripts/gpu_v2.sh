#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gpu_diag.py bf16 > gpurun_out/diag_v2.log 2>&1; echo "diag exit $?"; grep "paths=" gpurun_out/diag_v2.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^FAILED|^E   Assert" gpurun_out/pytest_gpu.log | head
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench exit $?"; python -c "
import json;d=json.load(open('gpurun_out/bench_v2.json'));print('v2 ms/step',d['ms_per_step'],'G7 ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'])"
WS_DISABLE_TC_V2=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; python -c "
import json;d=json.load(open('gpurun_out/bench_v1.json'));print('v1 ms/step',d['ms_per_step'],'G7 ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'])"
tail -3 gpurun_out/bench_v2.err
