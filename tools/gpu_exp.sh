#!/bin/bash
# one-off experiment driver (edited per experiment)
cd /root/repo
FN2_TC_KD=4 timeout 600 python tools/tc_conv_debug.py 2>&1 | grep "tc err" | cut -c1-120
for cl in 1 2 4; do
  echo "== CL=$cl"
  FN2_TC_CL=$cl FN2_TC_DBG=16 FN2_TC_KD=4 timeout 300 python tools/tc_time.py 2>&1 | grep "DBG="
done
