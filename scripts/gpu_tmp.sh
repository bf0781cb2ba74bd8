#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x --tb=short 2>&1 | tail -3
for sw in 1 0; do
  if [ $sw = 0 ]; then export WS_WGRAD_NO_SWAPXY=1; fi
  timeout 300 python bench.py --quick > gpurun_out/r02_bench_g$sw.json 2> gpurun_out/r02_bench_g$sw.err; echo "bench exit $?"
  python -c "
import json;d=json.load(open('gpurun_out/r02_bench_g$sw.json'));print('swapxy $sw', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline_step']['frac'], d['config']['per_step_ms']['resident'])"
done
