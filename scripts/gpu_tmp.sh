#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x --tb=short 2>&1 | tail -3
for lean in 1 0; do
  echo "== lean=$lean"
  for l in g7 g5 dg rdb; do
    WS_TC2_LEAN=$lean timeout 120 python scripts/prof_conv.py 5 $l 2>&1 | tail -1
  done
done
timeout 300 python bench.py --quick > gpurun_out/r02_bench_m.json 2> gpurun_out/r02_bench_m.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_m.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline_step']['frac'], d['gpu_launches'], d['config']['per_step_ms']['resident'])"
