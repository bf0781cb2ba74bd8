#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x --tb=short -k "two_rank" > gpurun_out/r02_pytest_2gpu_b.log 2>&1; echo "pytest exit $?"; grep DDP_RESULT gpurun_out/r02_pytest_2gpu_b.log; tail -2 gpurun_out/r02_pytest_2gpu_b.log
for mb in 64 256; do
  WINDSR_BUCKET_MB=$mb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --quick > gpurun_out/r02_bench_2gpu_mb$mb.json 2> gpurun_out/r02_bench_2gpu_mb$mb.err; echo "bench exit $?"
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_2gpu_mb$mb.json') if l.startswith('{')][-1]);print('bucket MB $mb', d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['config']['per_step_ms']['resident'])"
done
