#!/bin/bash
# 2-GPU sanity pass: the NCCL parity test and the weak-scaling bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x --tb=short -k "two_rank" > gpurun_out/r02_pytest_2gpu_final.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r02_pytest_2gpu_final.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --quick > gpurun_out/r02_bench_2gpu_final.json 2> gpurun_out/r02_bench_2gpu_final.err; echo "bench exit $?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_2gpu_final.json') if l.startswith('{')][-1]);print('2 GPUs', d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['config']['per_step_ms']['resident'])"
