#!/bin/bash
cd /root/repo
for lib in libfn2.so libfn2_base.so; do for kdw in 4 8; do
echo "== $lib KDW=$kdw"
FN2_LIB=/root/repo/flownet2_b200/$lib FN2_TC_KDW=$kdw timeout 300 python tools/tc_time.py 2>&1 | grep "DBG=" | cut -c1-75
done; done
