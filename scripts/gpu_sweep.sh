#!/bin/bash
for cfg in "4 6" "1 2" "1 6" "4 2" "2 6" "2 3"; do
  set -- $cfg
  for L in g7 g8 g5 rdb; do
    echo -n "ASPLIT=$1 ABUFS=$2  "; WS_TC2_ASPLIT=$1 WS_TC2_ABUFS=$2 python scripts/prof_conv.py 20 $L 2>&1 | tail -1
  done
done
