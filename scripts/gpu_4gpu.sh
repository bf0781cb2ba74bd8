#!/bin/bash
# 4-GPU sanity pass: the weak-scaling bench line
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --quick > gpurun_out/r02_bench_4gpu_final.json 2> gpurun_out/r02_bench_4gpu_final.err; echo "bench exit $?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_4gpu_final.json') if l.startswith('{')][-1]);print('4 GPUs', d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['config']['per_step_ms']['resident'])"
