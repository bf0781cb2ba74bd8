#!/bin/bash
for L in dg lff rdb0 rdb g5 g7; do python scripts/prof_conv.py 30 $L 2>&1 | tail -1; done
