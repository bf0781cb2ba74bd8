#!/bin/bash
for cfg in "4,4 6" "4,3 2" "4,3 3" "4,2 2" "4,2 3" "8,2 2" "8,1 3" "4,6 3" "4,8 2"; do
  set -- $cfg
  echo -n "FORCE=$1 ABUFS=$2  "; WS_TC2_FORCE=$1 WS_TC2_ABUFS=$2 python scripts/prof_conv.py 50 rdb 2>&1 | tail -1
done
