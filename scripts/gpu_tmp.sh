#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x --tb=short 2>&1 | tail -3
WS_RDB_DEBUG_TIMES=1 timeout 200 python scripts/prof_rdb.py 2>&1 | grep -A1 "rdb_fwd_persist\|rdb_bwd_persist\|default" | grep -v "^--" | head -8
timeout 300 python bench.py --quick > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_i.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline_step']['frac'], d['config']['per_step_ms']['resident'])"
