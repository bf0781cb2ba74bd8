#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_round_b.log
: > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x --tb=short >> $L 2>&1; echo "pytest exit $?" >> $L
for cfg in "1 1" "3 1" "3 4"; do
  set -- $cfg
  echo "== rdb persist w_group=$1 n_iss=$2" >> $L
  WS_RDB_WGROUP=$1 WS_RDB_NISS=$2 timeout 200 python scripts/prof_rdb.py 2>&1 | tail -4 >> $L
done
timeout 300 python bench.py --quick > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; echo "bench exit $?" >> $L
cat gpurun_out/r02_bench_b.json >> $L
tail -32 $L
