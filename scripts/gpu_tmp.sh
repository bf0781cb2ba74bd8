#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -q -x --tb=short 2>&1 | tail -3
timeout 600 python scripts/layer_sweep.py gpurun_out/r02_layer_sweep_v2.md 20 > gpurun_out/r02_layer_sweep_v2.log 2>&1; echo "sweep exit $?"
grep "^| D" gpurun_out/r02_layer_sweep_v2.md | cut -d'|' -f2,5-7,9-12
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_n.json 2> gpurun_out/r02_bench_n.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_n.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['config'].get('full_gan'))"
