#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_round_f.log
: > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x --tb=short >> $L 2>&1; echo "pytest exit $?" >> $L
for xh in 0 1; do
  echo "== XHALO=$xh" >> $L
  WS_WGRAD_XHALO=$xh timeout 120 python scripts/prof_conv.py 5 g7 wgrad 2>&1 | tail -1 >> $L
  WS_WGRAD_XHALO=$xh timeout 120 python scripts/prof_conv.py 5 g5 wgrad 2>&1 | tail -1 >> $L
  WS_WGRAD_XHALO=$xh timeout 120 python scripts/prof_trunk_wgrad.py "" 2>&1 | tail -2 >> $L
done
timeout 300 python bench.py --quick > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; echo "bench exit $?" >> $L
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_f.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline_step']['frac'], d['config']['per_step_ms'])" >> $L
tail -22 $L
