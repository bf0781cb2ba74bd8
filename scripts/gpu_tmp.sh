#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x --tb=short 2>&1 | tail -3
timeout 300 python bench.py --quick > gpurun_out/r02_bench_k.json 2> gpurun_out/r02_bench_k.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline_step']['frac'], d['gpu_launches'], d['config']['per_step_ms']['resident'])"
WS_XYUNFOLD_NO_TILE=1 timeout 300 python bench.py --quick > gpurun_out/r02_bench_k2.json 2> gpurun_out/r02_bench_k2.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_k2.json'));print('no tile', d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['per_step_ms']['resident'])"
