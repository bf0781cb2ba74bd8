#!/bin/bash
# one GPU round-trip: tests, bench line, per-kernel breakdown
timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_latest.json | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'G7',d['roofline']['kernel_ms'],d['gpu_launches'])"
python scripts/prof_step.py 2>&1 | tail -${1:-40}
