#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_tc2_dbg.log
: > $L
./scripts/micro/mma_rate2 >> $L 2>&1
for layer in g7 g5; do
  for cm in 0 1; do
    echo "== $layer chunk_major=$cm (plain)" >> $L
    WS_TC2_CHUNK_MAJOR=$cm timeout 120 python scripts/prof_conv.py 5 $layer >> $L 2>&1
    echo "== $layer chunk_major=$cm (debug timers)" >> $L
    WS_TC2_CHUNK_MAJOR=$cm WS_TC2_DEBUG_TIMES=1 timeout 120 python scripts/prof_conv.py 1 $layer 2>&1 | tail -3 >> $L
  done
done
# tile-shape alternatives for hr_convs.0 (by,tx) with 2 / 3 halo buffers
for f in "4,6" "4,9" "8,3" "8,4" "2,12" "4,3"; do
  echo "== g7 force $f" >> $L
  WS_TC2_FORCE=$f timeout 120 python scripts/prof_conv.py 5 g7 >> $L 2>&1
  WS_TC2_FORCE=$f WS_TC2_DEBUG_TIMES=1 timeout 120 python scripts/prof_conv.py 1 g7 2>&1 | tail -2 >> $L
done
cat $L
