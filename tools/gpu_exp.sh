#!/bin/bash
cd /root/repo
timeout 300 python tools/tc_time.py 2>&1 | grep "DBG="
FN2_TC_DBG=16 timeout 300 python tools/tc_time.py 2>&1 | grep -v "^$" | head -2
