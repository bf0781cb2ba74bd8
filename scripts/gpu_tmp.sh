#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x --tb=short 2>&1 | tail -3
WS_DISABLE_PDL=1 TIMELINE_NAME=r02_timeline2_nopdl.csv timeout 300 python scripts/prof_step.py g > gpurun_out/r02_prof_step2_nopdl.log 2>&1; echo "exit $?"
timeout 300 python bench.py --quick > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_e.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline_step']['frac'], d['config']['per_step_ms'])"
