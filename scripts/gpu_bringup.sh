#!/bin/bash
# First-contact run on the GPU box: kernel families one process at a time so a trap cannot hide other results.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python scripts/gpu_diag.py fp32 > gpurun_out/diag_fp32.log 2>&1; echo "fp32 exit $?" >> gpurun_out/summary.txt
for L in G4_lr G7_hr0 G3_lff G2_rdb0 G8_hr2 D2_strided D10_halve_z; do
  timeout 300 python scripts/gpu_diag.py bf16 $L > gpurun_out/diag_bf16_$L.log 2>&1; echo "bf16 $L exit $?" >> gpurun_out/summary.txt
done
timeout 600 python scripts/gpu_diag.py bf16 > gpurun_out/diag_bf16_all.log 2>&1; echo "bf16 all exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -h "paths=" gpurun_out/diag_fp32.log gpurun_out/diag_bf16_all.log | tail -40
