#!/bin/bash
# Round-2 evidence pass: full GPU test suite, smoke, layer sweep, parity table, ncu launch list + --set full captures of
# the top kernels, final bench line.  Everything comes back in gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed" gpurun_out/r02_pytest_final.log | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r02_smoke.log
timeout 600 python scripts/layer_sweep.py gpurun_out/r02_layer_sweep.md 20 > gpurun_out/r02_layer_sweep.log 2>&1; echo "sweep exit $?"
timeout 900 python scripts/parity_table.py --out gpurun_out/r02_parity_table_v2.md > gpurun_out/r02_parity_v2.log 2>&1; echo "parity exit $?"
# launch list of one steady-state step (eager; only after the same command exited 0 without ncu)
WINDSR_CUDA_GRAPH=0 timeout 300 python scripts/step_once.py 4 > gpurun_out/r02_step_once_plain.log 2>&1; echo "step_once exit $?"
WINDSR_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches_step_v2.csv python scripts/step_once.py 4 > gpurun_out/r02_step_once_ncu.log 2>&1; echo "ncu launches exit $?"
# --set full captures (one kernel each)
timeout 200 python scripts/prof_conv.py 2 g7 > gpurun_out/r02_plain_g7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc2 -s 1 -c 1 -f -o gpurun_out/r02_prof_g7_fwd_v2 \
  python scripts/prof_conv.py 2 g7 > gpurun_out/r02_ncu_g7_v2.log 2>&1; echo "ncu g7 exit $?"
timeout 200 python scripts/prof_conv.py 2 g7 wgrad > gpurun_out/r02_plain_g7w.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_g7_wgrad_v2 \
  python scripts/prof_conv.py 2 g7 wgrad > gpurun_out/r02_ncu_g7w_v2.log 2>&1; echo "ncu g7 wgrad exit $?"
timeout 200 python scripts/prof_rdb.py > gpurun_out/r02_plain_rdb.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rdb_.*_persist -s 6 -c 2 -f -o gpurun_out/r02_prof_rdb_persist_v2 \
  python scripts/prof_rdb.py > gpurun_out/r02_ncu_rdb_v2.log 2>&1; echo "ncu rdb exit $?"
timeout 200 python scripts/prof_trunk_wgrad.py "" > gpurun_out/r02_plain_trunkw.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02_prof_trunk_wgrad \
  python scripts/prof_trunk_wgrad.py "" > gpurun_out/r02_ncu_trunkw.log 2>&1; echo "ncu trunk wgrad exit $?"
# final bench line (all legs)
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit $?"
tail -c 1500 gpurun_out/r02_bench_final.json
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref bench exit $?"
tail -c 600 gpurun_out/r02_bench_reference.json
ls -la gpurun_out/*.ncu-rep | tail -5
