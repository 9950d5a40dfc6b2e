#!/bin/bash
cd /root/repo
timeout 600 python tools/tc_conv_debug.py 2>&1 | grep "tc err\|rror" | cut -c1-125
timeout 300 python tools/net_err.py FlowNet2 256 128 2>&1 | tail -1
FN2_TC_KD=2 timeout 300 python tools/net_err.py FlowNet2 256 128 2>&1 | tail -1
FN2_TC_KD=2 FN2_TC_KDW=4 timeout 300 python tools/net_err.py FlowNet2 256 128 2>&1 | tail -1
timeout 300 python tools/tc_time.py 2>&1 | grep "DBG="
