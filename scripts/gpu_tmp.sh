#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x --tb=short 2>&1 | tail -3
for l in g5 g7 g8; do
  echo "== $l"
  timeout 120 python scripts/prof_conv.py 5 $l 2>&1 | tail -1
  WS_TC2_DEBUG_TIMES=1 timeout 120 python scripts/prof_conv.py 1 $l 2>&1 | grep "tc2 dbg" | tail -1 | sed 's/.*accumulators/accumulators/'
done
timeout 300 python bench.py --quick > gpurun_out/r02_bench_l.json 2> gpurun_out/r02_bench_l.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_l.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline_step']['frac'], d['gpu_launches'], d['config']['per_step_ms']['resident'])"
