#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_round_a.log
: > $L
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x --tb=short >> $L 2>&1; echo "pytest exit $?" >> $L
for ni in 1 4; do
  echo "== g7 wgrad n_iss<=$ni" >> $L
  WS_WGRAD_NISS=$ni timeout 120 python scripts/prof_conv.py 5 g7 wgrad 2>&1 | tail -1 >> $L
  echo "== g5 wgrad n_iss<=$ni" >> $L
  WS_WGRAD_NISS=$ni timeout 120 python scripts/prof_conv.py 5 g5 wgrad 2>&1 | tail -1 >> $L
  echo "== trunk wgrad n_iss<=$ni" >> $L
  WS_WGRAD_NISS=$ni timeout 120 python scripts/prof_trunk_wgrad.py "" 4,1 2>&1 | tail -3 >> $L
done
timeout 300 python bench.py --quick > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench exit $?" >> $L
cat gpurun_out/r02_bench_a.json >> $L
tail -30 $L
