#!/bin/bash
# tile-shape sweep of the trunk convs on the halo kernel (WS_TC2_FORCE="by,tx")
for L in rdb rdb0 rdbx rdb0x dg dg0 lff; do
  echo "== $L default"; python scripts/prof_conv.py 200 $L | tail -1
  for F in 4,1 8,1 16,1 4,2 8,2 16,2 4,4 8,4 2,4 2,8 4,8; do
    echo -n "   force $F: "; WS_TC2_FORCE=$F python scripts/prof_conv.py 200 $L 2>&1 | tail -1
  done
done
