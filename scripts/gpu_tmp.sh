#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_round_c.log
: > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x --tb=short >> $L 2>&1; echo "pytest exit $?" >> $L
for e in 0 1; do
  echo "== rdb persist NO_EARLY=$e" >> $L
  WS_RDB_DEBUG_TIMES=1 WS_RDB_NO_EARLY=$e timeout 200 python scripts/prof_rdb.py 2>&1 | grep -A1 "rdb_fwd_persist\|default" | grep -v "^--" | head -4 >> $L
done
echo "== tap split" >> $L
WS_RDB_TAP_SPLIT=1 timeout 200 python scripts/prof_rdb.py 2>&1 | tail -1 >> $L
timeout 300 python bench.py --quick > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench exit $?" >> $L
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_c.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline_step']['frac'])" >> $L
tail -25 $L
