#!/bin/bash
# end-of-round pass: full GPU suite, smoke, the final bench lines and the launch list of one step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed" gpurun_out/r02_pytest_final.log | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r02_smoke.log
timeout 400 python bench.py --config upscale16 --quick > gpurun_out/r02_bench_upscale16.json 2> gpurun_out/r02_bench_upscale16.err; echo "upscale16 bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_upscale16.json'));print('upscale16', d['ms_per_step'], d['e2e']['ms_per_step'], d['value'])"
timeout 400 python bench.py --dtype tf32 --quick > gpurun_out/r02_bench_tf32_v2.json 2> gpurun_out/r02_bench_tf32_v2.err; echo "tf32 bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_tf32_v2.json'));print('tf32', d['ms_per_step'], d['e2e']['ms_per_step'], d['value'])"
WINDSR_CUDA_GRAPH=0 timeout 300 python scripts/step_once.py 4 > gpurun_out/r02_step_once_plain.log 2>&1; echo "step_once exit $?"
WINDSR_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches_step_v3.csv python scripts/step_once.py 4 > gpurun_out/r02_step_once_ncu.log 2>&1; echo "ncu launches exit $?"
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_final.json'));print('final', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline_step']['frac'], d['config'].get('full_gan'), d['config'].get('inference_b1'))"
