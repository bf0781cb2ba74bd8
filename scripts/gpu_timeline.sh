#!/bin/bash
mkdir -p gpurun_out
WS_DISABLE_PDL=1 TIMELINE_NAME=r02_timeline_nopdl.csv timeout 300 python scripts/prof_step.py g > gpurun_out/r02_prof_step_nopdl.log 2>&1; echo "exit $?"
TIMELINE_NAME=r02_timeline_pdl.csv timeout 300 python scripts/prof_step.py g > gpurun_out/r02_prof_step_pdl.log 2>&1; echo "exit $?"
WINDSR_TRUNK_BATCH=0 WS_DISABLE_PDL=1 TIMELINE_NAME=r02_timeline_nopdl_nobatch.csv timeout 300 python scripts/prof_step.py g > gpurun_out/r02_prof_step_nopdl_nobatch.log 2>&1; echo "exit $?"
head -8 gpurun_out/r02_prof_step_pdl.log
